"""bandwidthTest-style measurement of the host limit behind the multi-GPU `e2e` numbers: every rank copies a pinned
host buffer to its GPU (and back) concurrently; rank 0 prints per-rank and aggregate GB/s.  Launch with torchrun.
usage: python -m torch.distributed.run --nproc-per-node N tools/h2d_bandwidth.py [GiB=2]"""
import os
import sys
import time

import torch
import torch.distributed as dist

gib = float(sys.argv[1]) if len(sys.argv) > 1 else 2.0
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = int(gib * (1 << 30)) // 8
h = torch.empty(n, dtype=torch.float64).pin_memory()
h.fill_(1.0)
d = torch.empty(n, dtype=torch.float64, device="cuda")
out = {}
for name, (dst, src) in {"h2d": (d, h), "d2h": (h, d)}.items():
    for _ in range(2):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 8
    for _ in range(reps):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    out[name] = reps * n * 8 / dt / 1e9
t = torch.tensor([out["h2d"], out["d2h"]], dtype=torch.float64, device="cuda")
if world > 1:
    tl = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(tl, t)
else:
    tl = [t]
if rank == 0:
    h2d = [float(x[0]) for x in tl]; d2h = [float(x[1]) for x in tl]
    print(f"H2D_BANDWIDTH world={world} pinned {gib:.1f} GiB per rank: H2D per rank {min(h2d):.1f}-{max(h2d):.1f} GB/s, aggregate {sum(h2d):.1f} GB/s; "
          f"D2H per rank {min(d2h):.1f}-{max(d2h):.1f} GB/s, aggregate {sum(d2h):.1f} GB/s (concurrent within a direction)")
if world > 1:
    dist.destroy_process_group()
