#!/usr/bin/env python
"""bench.py -- cells/s condensed+assembled (FP64), headline: the 3-D HDG k=2 configuration (BASELINE.json C3).

A "step" is one pass of the hot path over the synthetic mesh of the configuration: static condensation of every cell
(A_K,b_K)->(S_K,g_K) followed by the numeric skeleton assembly into the cached CSC pattern.  Inputs are resident in HBM
before the timed region (`value`); `e2e` repeats the same metric through the C-ABI call `ghb_condense_assemble_f64` with
HOST buffers (H2D of the records and D2H of nzval/rhs inside the timed region) on a bounded sample of the same
workload -- pinned caller memory (`e2e`) and pageable caller memory staged by the library (`e2e_pageable`).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config C3|C1|C2k1|C2k2|C2k3|C4|C5]
                  [--scaling strong|weak] [--dims ...]

Multi-GPU (C3): one process per GPU (torchrun); the cells are partitioned in z-slabs.  `--scaling strong` (default, what
BASELINE.json names: "2M cells, sharded across 1/2/4/8 B200") divides the 128^3 mesh over the ranks; `--scaling weak`
gives every rank its own 128^3 slab.  The cut-plane exchange (NCCL send/recv through the C ABI, ghb_exchange_cut_plane_f64)
is inside the timed step; the backward step with the lambda all-gather (ghb_allgather_lambda_f64) is timed on its own.
Timing: CUDA events, barrier + synchronize on both sides, max over ranks.

Configurations other than C3 run on one GPU.  Where records + S + CSC of the named mesh exceed HBM (C2 k=3, C4, C5) the
mesh is processed in sequential z-slabs of equal size; every slab is assembled as a mesh of its own (its cut planes are
Dirichlet boundaries, 1-2 % of the facets) and its records are generated on the device before its timed region.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

RTH = np.array([[1, 1, 1], [1, 0, 0], [1, 0, 0]], bool)
_HENCKY = np.ones((8, 8), bool)
for _i, _j in [(0, 1), (1, 0), (2, 4), (4, 2), (6, 7), (7, 6), (0, 7), (7, 0)]:
    _HENCKY[_i, _j] = False
# BASELINE.json configs: fields, touched mask, mesh, dofs per facet of the skeleton space
CONFIGS = {
    "C1": dict(label="C1 Darcy HDG k=1 2-D quad Cartesian", ndofs=[6, 1, 8], touched=np.ones((3, 3), bool),
               interior=[1, 2], boundary=[3], dims=(32, 32), ndofs_f=2),
    "C2k1": dict(label="C2 Darcy RT-H k=1 2-D quads", ndofs=[12, 4, 8], touched=RTH, interior=[1, 2], boundary=[3],
                 dims=(2048, 2048), ndofs_f=2),
    "C2k2": dict(label="C2 Darcy RT-H k=2 2-D quads", ndofs=[24, 9, 12], touched=RTH, interior=[1, 2], boundary=[3],
                 dims=(2048, 2048), ndofs_f=3),
    "C2k3": dict(label="C2 Darcy RT-H k=3 2-D quads", ndofs=[40, 16, 16], touched=RTH, interior=[1, 2], boundary=[3],
                 dims=(2048, 2048), ndofs_f=4),
    "C3": dict(label="C3 Darcy HDG k=2 3-D hex Cartesian", ndofs=[30, 4, 36], touched=np.ones((3, 3), bool),
               interior=[1, 2], boundary=[3], dims=(128, 128, 128), ndofs_f=6),
    "C4": dict(label="C4 elasticity HDG k=2 3-D hexes", ndofs=[60, 60, 108], touched=np.ones((3, 3), bool),
               interior=[1, 2], boundary=[3], dims=(100, 100, 100), ndofs_f=18),
    # two skeleton fields (3 + 9 dofs per facet): the cell ids of ONE 12-dof field stand in for them (same pattern up to
    # a relabelling of the dofs; multi-field ids are a parity test, tests/test_gpu_parity.py)
    "C5": dict(label="C5 Hencky HDG k=1 3-D hexes", ndofs=[12, 12, 4, 24, 24, 30, 18, 54], touched=_HENCKY,
               interior=[1, 2, 3, 4, 5, 6], boundary=[7, 8], dims=(200, 200, 200), ndofs_f=12),
}
METRIC = "cells/s condensed+assembled (FP64, 3D HDG k=2)"
FP64_PEAK_TFLOPS = 37.1      # measured DMMA peak of this pool's B200 (profiles/r01_ubench_fp64.txt); tcgen05 has no FP64 kind
HBM_BUDGET = 165e9           # bytes of a GPU's 180 GB a step may keep resident

# dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel from the committed `ncu --set full` capture, per cell
NCU_DRAM_BYTES_PER_CELL = {"cw_34_36": 57108}
NCU_TRAFFIC_SOURCE = "profiles/r02_cw_summary.md (ncu --set full, 131072-cell launch, scaled per cell)"


def shape_of(cfg):
    nd, t = cfg["ndofs"], cfg["touched"]
    n_i = sum(nd[f - 1] for f in cfg["interior"]); n_b = sum(nd[f - 1] for f in cfg["boundary"])
    lenA = int(sum(nd[i] * nd[j] for i in range(len(nd)) for j in range(len(nd)) if t[i, j]))
    return n_i, n_b, lenA, int(sum(nd))


def algorithmic_bytes_per_cell(cfg):
    """B_cond = 8[(touched entries of A_K + b_K) + n_b^2 + n_b]  (read the record once, write S_K,g_K once)."""
    n_i, n_b, lenA, lenb = shape_of(cfg)
    return 8 * (lenA + lenb + n_b * n_b + n_b)


def flops_per_cell(cfg):
    n_i, n_b, _, _ = shape_of(cfg)
    return (2.0 / 3.0) * n_i ** 3 + 2 * n_i ** 2 * (n_b + 1) + 2 * n_i * n_b * (n_b + 1)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.stop, self.t = index, [], threading.Event(), None

    def _run(self):
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop.wait(0.1)

    def __enter__(self):
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.t.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(r[3 + k].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "samples": len(self.rows), "power_w_max": max(float(r[2]) for r in self.rows)}


# --------------------------------------------------------------------------------------------------
# CPU arm: the oracle's C twin (condensation, the reference's LAPACK sequence per cell) + SciPy COO->CSC (the reference's
# serial sparse(I,J,V)) on a bounded sample of the workload, Philox records identical to the device generator
# --------------------------------------------------------------------------------------------------

def cpu_sample(cfg, sample_dims, steps, warmup):
    from oracle import oracle as o
    from oracle import oracle_c as oc
    plan = o.BlockPlan(cfg["ndofs"], cfg["touched"], cfg["interior"], cfg["boundary"])
    n_i, n_b, _, _ = shape_of(cfg)
    n = int(np.prod(sample_dims))
    A, b = o.synth_cell_records(plan, 0, n)                 # the same counter-based records the device arm generates
    cwf = cartesian_cwf_numpy(sample_dims)
    ids = facet_ids_numpy(cwf, sample_dims, cfg["ndofs_f"])
    nfree = int(ids.max())
    cores = oc.max_threads()
    split = {}

    def step(nthreads):
        t0 = time.perf_counter()
        S, g, info = oc.condense(plan, A, b, nthreads=nthreads)
        t1 = time.perf_counter()
        # Gridap's COO numeric loop + SparseArrays.sparse! restated in C (oracle_c.c), serial like the reference
        colptr, rowval, nzval, rhs = oc.assemble_coo_csc(S, g, ids, nfree)
        split["condense_ms"], split["assemble_ms"] = (t1 - t0) * 1e3, (time.perf_counter() - t1) * 1e3
        return (colptr, rowval, nzval), rhs

    for _ in range(warmup):
        step(0)
    t0 = time.perf_counter()
    for _ in range(steps):
        step(0)
    dt = (time.perf_counter() - t0) / steps
    all_split = dict(split)
    t0 = time.perf_counter()
    step(1)                                                  # the reference itself is serial: one core
    dt1 = time.perf_counter() - t0
    sample = (f"{n} cells ({'x'.join(map(str, sample_dims))} mesh), Philox records: C oracle (dgetrf/dgetrs/dgemm sequence per "
              f"cell, pthreads, {cores} threads) condensation {all_split['condense_ms']:.0f} ms + assembly "
              f"{all_split['assemble_ms']:.0f} ms (C restatement of Gridap's COO loop + SparseArrays.sparse!, serial like the "
              f"reference's sparse(I,J,V)); 1 thread: {n / dt1:.0f} cells/s; "
              f"Julia reference not runnable here (no julia binary; JULIA_NUM_THREADS=n/a, the reference is single-threaded)")
    return n / dt, cores, dt, sample, n / dt1


def cartesian_cwf_numpy(dims):
    """closed-form first-touch facet ids (vectorised twin of the device kernel; bench CPU arm only)."""
    D = len(dims)
    stride = np.concatenate([[1], np.cumprod(dims)])
    c = np.arange(stride[-1], dtype=np.int64)

    def own(cc, axis, side):
        run = D * cc
        for a in range(D):
            hi, rem = cc // stride[a + 1], cc % stride[a + 1]
            run = run + hi * stride[a] + np.minimum(rem, stride[a])
        res = np.zeros_like(cc)
        for a in range(D - 1, -1, -1):
            ia = (cc // stride[a]) % dims[a]
            run = run + (ia == 0)
            if a == axis and side == 0:
                res = run.copy()
            run = run + 1
            if a == axis and side == 1:
                res = run.copy()
        return res

    out = np.zeros((len(c), 2 * D), dtype=np.int64)
    for lf in range(2 * D):
        axis, side = D - 1 - lf // 2, lf % 2
        ia = (c // stride[axis]) % dims[axis]
        if side == 1:
            out[:, lf] = own(c, axis, 1)
        else:
            lowb = own(c, axis, 0)
            nb = own(np.maximum(c - stride[axis], 0), axis, 1)
            out[:, lf] = np.where(ia == 0, lowb, nb)
    return out


def facet_ids_numpy(cwf, dims, ndofs_f):
    nfacets = int(cwf.max())
    cnt = np.bincount(cwf.ravel() - 1, minlength=nfacets)
    isd = cnt == 1
    free_rank = np.cumsum(~isd)
    dir_rank = np.cumsum(isd)
    d = np.arange(1, ndofs_f + 1)[None, :]
    fid = np.where(isd[:, None], -((dir_rank[:, None] - 1) * ndofs_f + d), (free_rank[:, None] - 1) * ndofs_f + d)
    return fid[cwf - 1].reshape(cwf.shape[0], -1)


def workload_name(cfg, dims):
    """the one workload both arms report (the reference arm times a bounded sample of it)"""
    n_i, n_b, _, _ = shape_of(cfg)
    return (f"{cfg['label']} {'x'.join(map(str, dims))} cells (n_i={n_i}, n_b={n_b}), all boundary facets Dirichlet, "
            f"Philox synthetic records")


def cpu_sample_dims(cfg):
    """bounded CPU sample: about 10-30 s of host work"""
    n_i, n_b, _, _ = shape_of(cfg)
    if len(cfg["dims"]) == 2:
        return (min(cfg["dims"][0], 256), min(cfg["dims"][1], 128 if n_i < 20 else 64))
    return (32, 32, 24) if n_i < 64 else (12, 12, 8)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = CONFIGS[args.config]
    dims = tuple(args.dims) if args.dims else cfg["dims"]
    val, cores, dt, sample, val1 = cpu_sample(cfg, cpu_sample_dims(cfg), args.steps, max(args.warmup, 1))
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "cells/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(cfg, dims)},
            "cpu_baseline": {"value": val, "unit": "cells/s", "cores": cores, "kind": "port", "sample": sample,
                             "value_1core": val1},
            "e2e": {"value": val, "unit": "cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------

def bind_to_gpu_numa_node(local):
    """pin this process (and the pages it first-touches: pinned staging buffers) to the NUMA node of its GPU"""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        dom, rest = bus.split(":", 1)
        node = int(open(f"/sys/bus/pci/devices/{dom[-4:].lower()}:{rest.lower()}/numa_node").read())
        if node < 0:
            return None
        cpus = []
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus += list(range(int(a), int(b or a) + 1))
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


def slab_chunks(cfg, dims, world):
    """(z extent per rank, sequential sub-slabs per rank, z extent of a sub-slab): the largest equal sub-slabs whose
    records + S + CSC stay inside HBM_BUDGET"""
    n_i, n_b, lenA, lenb = shape_of(cfg)
    D = len(dims)
    nnz_cell = D * cfg["ndofs_f"] ** 2 * (4 * D - 1)
    per_cell = 8 * (lenA + lenb) + 8 * (n_b * n_b + n_b) + 18 * nnz_cell + 8 * n_b * 3 + 64
    zr = dims[-1] // world
    layer = int(np.prod(dims[:-1]))
    zc = zr
    while zc > 1 and per_cell * layer * zc > HBM_BUDGET:
        zc -= 1
        while zr % zc:
            zc -= 1
    return zr, zr // zc, zc


def run_ours(args):
    import torch
    import torch.distributed as dist
    import gridaphybrid_b200 as gh
    from gridaphybrid_b200.distributed import SlabAssembler

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world} (launch with torchrun for N>1)"
    numa = bind_to_gpu_numa_node(local)
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = gh.Context(local)
    gh.set_default_context(ctx)
    ctx.use_torch_stream()
    dev = torch.device("cuda", local)
    if world > 1:
        ctx.comm_init_from_torch()          # NCCL communicator inside the library; torch only ships the 128-byte id

    cfg = CONFIGS[args.config]
    n_i, n_b, lenA, lenb = shape_of(cfg)
    base = tuple(args.dims) if args.dims else cfg["dims"]
    D = len(base)
    if world > 1:
        assert args.config == "C3", "multi-GPU runs are the C3 configuration (BASELINE.json config 3)"
    gdims = base[:-1] + (base[-1] * world,) if args.scaling == "weak" else base
    assert gdims[-1] % world == 0, "the slowest axis must be divisible by the number of GPUs"
    zr, nchunk, zc = slab_chunks(cfg, gdims, world)
    layer = int(np.prod(gdims[:-1]))
    ncells_rank = layer * zr
    plan = ctx.plan_blocks(cfg["ndofs"], cfg["touched"], cfg["interior"], cfg["boundary"])
    if nchunk == 1:
        slab = SlabAssembler(ctx, gdims, cfg["ndofs_f"], rank, world)
        cdims = gdims[:-1] + (zr,)
    else:
        assert world == 1
        cdims = gdims[:-1] + (zc,)
        slab = SlabAssembler(ctx, cdims, cfg["ndofs_f"], 0, 1)       # every sub-slab is a mesh of its own, same pattern
    ncells = layer * (zr if nchunk == 1 else zc)                      # cells of one timed pass
    A = torch.empty((ncells, plan.lenA), dtype=torch.float64, device=dev)
    b = torch.empty((ncells, plan.lenb), dtype=torch.float64, device=dev)
    S = torch.empty((ncells, n_b * n_b), dtype=torch.float64, device=dev)
    g = torch.empty((ncells, n_b), dtype=torch.float64, device=dev)
    info = torch.empty((ncells,), dtype=torch.int32, device=dev)
    nzval = torch.empty(slab.nnz, dtype=torch.float64, device=dev)
    rhs = torch.empty(slab.nrows_local, dtype=torch.float64, device=dev)
    cell_start = rank * ncells_rank
    ctx.synth_fill(plan, cell_start, ncells, A, b)
    torch.cuda.synchronize()

    ev = lambda: torch.cuda.Event(enable_timing=True)
    cond_ms, step_ms = [], []

    fused = bool(args.fused) and plan.kernel_name.startswith("cw")

    def one_pass(timed):
        e0, e1, e2, ea = ev(), ev(), ev(), ev()
        ea.record()
        if fused:
            # one kernel condenses and scatters S_K into the zeroed nzval (no S round trip, no gather kernel); the events
            # e0..e1 bracket that kernel alone, ea..e2 the whole pass (memset of nzval, pack + exchange, ghosts, rhs)
            slab.condense_assemble(plan, A, b, S, g, info, nzval, rhs, events=(e0, e1))
        else:
            e0.record()
            ctx.condense(plan, ncells, A, b, S, g, info)
            e1.record()
            slab.assemble(S, g, nzval, rhs)
        e2.record()
        if timed:
            cond_ms.append((e0, e1)); step_ms.append((ea, e2))

    def step(timed):
        for k in range(nchunk):
            if nchunk > 1:
                ctx.synth_fill(plan, cell_start + k * ncells, ncells, A, b)     # outside the events of the pass
            one_pass(timed)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for _ in range(args.warmup):
        step(False)
    barrier()
    l0 = ctx.launch_count
    with ClockSampler(local) as clk:
        t0, t1 = ev(), ev()
        t0.record()
        for _ in range(args.steps):
            step(True)
        t1.record()
        barrier()
    launches = ctx.launch_count - l0
    if nchunk == 1:
        ms = max_over_ranks(t0.elapsed_time(t1) / args.steps)
    else:   # the record generation of a sub-slab sits between the timed passes: sum of the passes' own events
        ms = sum(a.elapsed_time(bb) for a, bb in step_ms) / args.steps
    assert int(info.abs().sum().item()) == 0, "a cell failed to factorise"
    kernel_ms = float(np.mean([a.elapsed_time(bb) for a, bb in cond_ms]))
    total_cells = ncells_rank * world
    value = total_cells / (ms * 1e-3)
    launches -= (nchunk if nchunk > 1 else 0) * args.steps           # synth_fill launches are not the path

    # ---- the other assembly variant, measured next to the headline (same records, same pattern) ----
    other = None
    if plan.kernel_name.startswith("cw") and nchunk == 1:
        def other_pass():
            if fused:
                ctx.condense(plan, ncells, A, b, S, g, info)
                slab.assemble(S, g, nzval, rhs)
            else:
                slab.condense_assemble(plan, A, b, S, g, info, nzval, rhs)
        for _ in range(2):
            other_pass()
        barrier()
        o0, o1 = ev(), ev()
        o0.record()
        for _ in range(max(2, args.steps // 2)):
            other_pass()
        o1.record()
        barrier()
        oms = max_over_ranks(o0.elapsed_time(o1) / max(2, args.steps // 2))
        other = {"value": total_cells / (oms * 1e-3), "unit": "cells/s", "ms_per_step": oms,
                 "variant": "condense, then gather_nzval (two kernels)" if fused else
                            "one kernel: the condensation scatters S_K into the zeroed nzval with FP64 atomics (<= 2 contributions "
                            "per entry: bit-equal to the gather, tests/test_gpu_parity.py::test_fused_scatter_assembly_bit_equal_to_gather)"}
        if not fused:
            ctx.condense(plan, ncells, A, b, S, g, info)      # leave all of S_K behind for the legs below

    # ---- launch-bound configurations (records smaller than L2, e.g. C1: 1 024 cells): the step captured in a CUDA graph
    # (the C-ABI calls with device pointers are pure launch sequences on the context's stream) and replayed
    graph_leg = None
    if world == 1 and nchunk == 1 and ncells * (lenA + lenb) * 8 <= 256e6:
        gph = torch.cuda.CUDAGraph()
        torch.cuda.synchronize()
        with torch.cuda.graph(gph):
            ctx.use_torch_stream()                              # the capture stream
            ctx.condense(plan, ncells, A, b, S, g, info)
            slab.assemble(S, g, nzval, rhs)
        ctx.use_torch_stream()
        nrep = max(20, args.steps)
        for _ in range(3):
            gph.replay()
        barrier()
        q0, q1 = ev(), ev()
        q0.record()
        for _ in range(nrep):
            gph.replay()
        q1.record()
        barrier()
        gms = q0.elapsed_time(q1) / nrep
        graph_leg = {"value": total_cells / (gms * 1e-3), "unit": "cells/s", "ms_per_step": gms,
                     "note": "the same two-kernel step captured once in a CUDA graph and replayed (launch-bound configuration)"}
        del gph

    # ---- N > 1: the assembled system against a local property of the condensed cells (no second copy of the mesh needed):
    # sum(nzval) over all ranks == sum over all cells of the free-free entries of S_K (cut-plane contributions included)
    check = None
    if world > 1:
        ctx.condense(plan, ncells, A, b, S, g, info)          # all of S_K (the fused step keeps only what it needs)
        ids = slab.cell_ids[:ncells]
        tot = torch.zeros((), dtype=torch.float64, device=dev)
        for c0 in range(0, ncells, 32768):
            m = (ids[c0:c0 + 32768] > 0).to(torch.float64)
            Sv = S[c0:c0 + 32768].view(-1, n_b, n_b)
            tot += torch.einsum("ci,cij,cj->", m, Sv.transpose(1, 2), m)
        gs = (g * (ids > 0)).sum()
        red = torch.stack([tot, nzval.sum(), gs, rhs.sum(), nzval.abs().sum()])
        dist.all_reduce(red)
        r = red.cpu().numpy()
        check = {"sum_nzval_rel_diff": float(abs(r[0] - r[1]) / r[4]), "sum_rhs_rel_diff": float(abs(r[2] - r[3]) / max(abs(r[2]), 1e-300)),
                 "what": "all-reduced sum(nzval) / sum(rhs) against the free-free entries of the local S_K / g_K"}
        assert check["sum_nzval_rel_diff"] < 1e-11, check

    # ---- backward map on the same records: lambda all-gather (C ABI, grouped ncclBroadcast) + back-substitution, timed
    # together; reported on its own
    backsub = None
    if nchunk == 1:
        L = slab.layout
        lam_owned = torch.randn(L.nrows_local, dtype=torch.float64, device=dev)    # this rank's range of the skeleton solution
        u = torch.empty((ncells, plan.n_i), dtype=torch.float64, device=dev)
        ids_local = slab.cell_ids[:ncells]

        def back():
            lam = slab.allgather_lambda(lam_owned)
            ctx.backsub(plan, ncells, A, b, lam, None, ids_local, u, info)

        for _ in range(2):
            back()
        barrier()
        b0, b1 = ev(), ev()
        b0.record()
        for _ in range(3):
            back()
        b1.record()
        barrier()
        back_ms = max_over_ranks(b0.elapsed_time(b1) / 3)
        backsub = {"value": total_cells / (back_ms * 1e-3), "unit": "cells/s", "ms": back_ms,
                   "note": "lambda all-gather (ghb_allgather_lambda_f64) + BackwardStaticCondensationMap on the same records "
                           "(LU recomputed, as the reference does)"}
        del lam_owned, u

    # ---- e2e: C-ABI with host buffers, bounded sample, copies inside the timed region ----
    e2e, e2e_pageable, cpu_base, e2e_affine = None, None, None, None
    sdims = tuple(args.e2e_dims) if args.e2e_dims else ((64, 64, 32) if D == 3 and n_i < 64 else ((24, 24, 16) if D == 3 else (512, 256)))
    sn = min(int(np.prod(sdims)), ncells)
    if sn == int(np.prod(sdims)):
        ssk = gh.CartesianSkeleton(sdims, ctx)
        sM = gh.FacetFESpace(ssk, cfg["ndofs_f"], ssk.facet_is_boundary())
        sass = gh.SparseMatrixAssembler(sM)
        _, _, snnz = sass.symbolic()
        hA = torch.empty((sn, plan.lenA), dtype=torch.float64).pin_memory()
        hb = torch.empty((sn, plan.lenb), dtype=torch.float64).pin_memory()
        hA.copy_(A[:sn]); hb.copy_(b[:sn])
        hz = torch.empty(snnz, dtype=torch.float64).pin_memory()
        hr = torch.empty(sass.nrows, dtype=torch.float64).pin_memory()
        hinfo = torch.empty(sn, dtype=torch.int32).pin_memory()
        h2d = int(sn * (plan.lenA + plan.lenb) * 8)
        d2h = int((snnz + sass.nrows) * 8 + sn * 4)

        def timed_e2e(Ah, bh, zh, rh, ih, reps):
            sass.select()
            for _ in range(2):
                ctx.condense_assemble(plan, sn, Ah, bh, None, zh, rh, ih)
            barrier()
            t0 = time.perf_counter()
            for _ in range(reps):
                ctx.condense_assemble(plan, sn, Ah, bh, None, zh, rh, ih)   # returns after the D2H completed
            torch.cuda.synchronize()
            return max_over_ranks((time.perf_counter() - t0) / reps)

        dt = timed_e2e(hA, hb, hz, hr, hinfo, args.steps)
        e2e = {"value": sn * world / dt, "unit": "cells/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "h2d_gbs_aggregate": h2d * world / dt / 1e9, "numa_node": numa,
               "sample": f"{sn} cells per GPU ({'x'.join(map(str, sdims))}) via ghb_condense_assemble_f64, pinned host buffers"}
        # pageable caller memory (what a Julia Array is): staged through the library's pinned double buffer by host threads
        pA, pb = hA.numpy().copy(), hb.numpy().copy()
        pz, pr, pi = np.empty(snnz), np.empty(sass.nrows), np.empty(sn, dtype=np.int32)
        dtp = timed_e2e(pA, pb, pz, pr, pi, max(2, args.steps // 2))
        assert np.array_equal(pz, hz.numpy()), "pageable and pinned paths disagree"
        e2e_pageable = {"value": sn * world / dtp, "unit": "cells/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "h2d_gbs_aggregate": h2d * world / dtp / 1e9,
                        "sample": "same sample, pageable numpy arrays in and out (pinned staging inside the library)"}
        # the same numpy arrays page-locked in place through the C ABI (ghb_host_register: what a Julia glue does once per
        # array): copied from at the PCIe rate like torch's pinned tensors; the registration itself is timed separately
        t0 = time.perf_counter()
        for arr in (pA, pb, pz, pr):
            ctx.host_register(arr)
        reg_s = time.perf_counter() - t0
        dtr = timed_e2e(pA, pb, pz, pr, pi, max(2, args.steps // 2))
        for arr in (pA, pb, pz, pr):
            ctx.host_unregister(arr)
        assert np.array_equal(pz, hz.numpy()), "registered and pinned paths disagree"
        e2e_pageable["registered_in_place"] = {"value": sn * world / dtr, "unit": "cells/s", "register_seconds_once": reg_s,
                                               "note": "the same numpy arrays after ghb_host_register (cudaHostRegister)"}
        # informational: past the PCIe ceiling of the record path -- for an affine family the host ships tables + per-cell
        # coefficient vectors (ghb_expand_records_f64 generates the records on the device) and gets the CSC values back
        e2e_affine = None
        if args.config == "C3" and world == 1:
            ntab = 7
            rng = np.random.default_rng(5)
            TAh = torch.as_tensor(np.concatenate([A[:1].cpu().numpy(), 1e-3 * rng.standard_normal((ntab - 1, plan.lenA))])).pin_memory()
            Tbh = torch.as_tensor(np.concatenate([b[:1].cpu().numpy(), 1e-3 * rng.standard_normal((ntab - 1, plan.lenb))])).pin_memory()
            coefh = gh.cartesian_coefficients(sdims, tuple(1.0 / d for d in sdims), dev).cpu().pin_memory()

            def affine_step():      # host tables + coefficients in, CSC values + rhs out (host): ONE library call, and for a
                ctx.condense_assemble_affine(plan, sn, ntab, TAh, Tbh, coefh, None, hz, hr, hinfo)   # cell-warp plan ONE kernel

            sass.select()
            for _ in range(2):
                affine_step()
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                affine_step()
            torch.cuda.synchronize()
            dta = (time.perf_counter() - t0) / args.steps
            e2e_affine = {"value": sn / dta, "unit": "cells/s",
                          "h2d_bytes_per_step": int((TAh.numel() + Tbh.numel() + coefh.numel()) * 8), "d2h_bytes_per_step": d2h,
                          "note": "affine record family (7 tables): tables + coefficient vectors from pinned host memory through "
                                  "ghb_condense_assemble_affine_f64 (records formed in the loader of the condensation kernel, "
                                  "S_K scattered into nzval), host nzval/rhs; D2H-bound"}
        del hA, hb, hz, hr, pA, pb, pz, pr

    # ---- informational: the same step with the records generated on the device from an affine family:
    # expand -> condense -> assemble from per-cell coefficient vectors; nothing of the size of the records crosses PCIe.
    # (Overwrites A, b: last leg.  The `e2e` key above stays the host-record path the contract asks for.)
    devgen = None
    if args.config == "C3" and nchunk == 1:
        ntab = 1 + 2 * 3
        rng = np.random.default_rng(3)
        A0r, b0r = torch.empty((1, plan.lenA), dtype=torch.float64, device=dev), torch.empty((1, plan.lenb), dtype=torch.float64, device=dev)
        ctx.synth_fill(plan, 0, 1, A0r, b0r)                      # the same base record on every rank
        TA = np.concatenate([A0r.cpu().numpy(), 1e-3 * rng.standard_normal((ntab - 1, plan.lenA))])
        Tb = np.concatenate([b0r.cpu().numpy(), 1e-3 * rng.standard_normal((ntab - 1, plan.lenb))])
        fam = gh.AffineRecordFamily(TA, Tb)
        coef = gh.cartesian_coefficients(gdims, tuple(1.0 / d for d in gdims), dev, cell_start=cell_start, ncells=ncells)

        def gen_step():          # records written to HBM by ghb_expand_records_f64, then the resident-record step
            fam.expand(ctx, plan, coef, A, b)
            if fused:
                slab.condense_assemble(plan, A, b, S, g, info, nzval, rhs)
            else:
                ctx.condense(plan, ncells, A, b, S, g, info)
                slab.assemble(S, g, nzval, rhs)

        def gen_fused_step():    # records formed in the loader of the condensation kernel: they never exist in HBM
            slab.condense_assemble_affine(plan, fam, coef, S, g, info, nzval, rhs)      # + cut-plane exchange at N > 1

        def timed3(f):
            f()
            barrier()
            g0, g1 = ev(), ev()
            g0.record()
            for _ in range(3):
                f()
            g1.record()
            barrier()
            return max_over_ranks(g0.elapsed_time(g1) / 3)

        gms2 = timed3(gen_step)
        assert int(info.abs().sum().item()) == 0
        ref_sum = (float(nzval.sum().item()), float(rhs.sum().item()))
        gms = timed3(gen_fused_step)
        assert int(info.abs().sum().item()) == 0
        same = ref_sum == (float(nzval.sum().item()), float(rhs.sum().item()))       # the two paths are bit-identical
        devgen = {"value": total_cells / (gms * 1e-3), "unit": "cells/s", "ms": gms, "h2d_bytes_per_step": int(coef.numel() * 8),
                  "note": "records of an affine family (7 tables) formed inside the condensation kernel "
                          "(ghb_condense_scatter_slab_affine_f64 + ghb_assemble_finish_slab_f64: TMA-staged table chunks, DMMA combination per batch of 8 cells, "
                          "scratch records in L2), condensed and assembled; coefficients counted as the host input",
                  "expand_then_condense": {"value": total_cells / (gms2 * 1e-3), "ms": gms2,
                                           "note": "records written to HBM by ghb_expand_records_f64 first (round-2 path)"},
                  "checksums_equal": bool(same)}
        # the backward map from the same coefficient vectors (ghb_backsub_affine_f64): with the step above the whole solve
        # runs without the 83 GB of records ever existing
        lam_b = torch.randn(slab.layout.nrows_local, dtype=torch.float64, device=dev)
        u_b = torch.empty((ncells, plan.n_i), dtype=torch.float64, device=dev)
        ids_b = slab.cell_ids[:ncells]
        bms = timed3(lambda: fam.backsub(ctx, plan, coef, slab.allgather_lambda(lam_b), None, ids_b, u_b, info))
        assert int(info.abs().sum().item()) == 0
        devgen["backsub"] = {"value": total_cells / (bms * 1e-3), "unit": "cells/s", "ms": bms,
                             "note": "BackwardStaticCondensationMap with the records formed in the loader (GEN + BACK kernel)"}
        del lam_b, u_b
        del coef

    if rank == 0:
        if world == 1 and not args.no_cpu:
            v, cores, _, sample, v1 = cpu_sample(cfg, cpu_sample_dims(cfg), 1, 1)
            cpu_base = {"value": v, "unit": "cells/s", "cores": cores, "kind": "port", "sample": sample, "value_1core": v1}
        peak, how = measured_peaks()
        B, F = algorithmic_bytes_per_cell(cfg), flops_per_cell(cfg)
        hbm_bound = peak * 1e9 / B                      # cells/s if the kernel moved its algorithmic bytes at the HBM peak
        fp_bound = FP64_PEAK_TFLOPS * 1e12 / F
        krate = ncells / (kernel_ms * 1e-3)
        if hbm_bound <= fp_bound:
            ach = B * krate / 1e9
            roof = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "peak_source": how}
        else:
            ach = F * krate / 1e12
            roof = {"bound": "tensor", "achieved": ach, "peak": FP64_PEAK_TFLOPS, "unit": "TFLOP/s", "frac": ach / FP64_PEAK_TFLOPS,
                    "peak_source": "FP64 DMMA peak measured on this pool (profiles/r01_ubench_fp64.txt); tcgen05 has no FP64 kind"}
        tr = NCU_DRAM_BYTES_PER_CELL.get(plan.kernel_name)
        roof.update({"traffic": tr * ncells if tr else None, "traffic_source": NCU_TRAFFIC_SOURCE if tr else None,
                     "kernel": plan.kernel_name, "kernel_ms": kernel_ms, "algorithmic_bytes_per_cell": B,
                     "flops_per_cell": F, "fp64_tflops": F * krate / 1e12, "hbm_gbs": B * krate / 1e9,
                     "bound_cells_per_s": min(hbm_bound, fp_bound), "kernel_cells_per_s": krate})
        line = {"metric": METRIC if args.config == "C3" else METRIC.replace("3D HDG k=2", args.config), "value": value,
                "unit": "cells/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": args.scaling,
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload_name(cfg, gdims), "config": args.config,
                           "cells_per_gpu": ncells_rank, "sequential_sub_slabs": nchunk,
                           "sub_slab": "x".join(map(str, cdims)), "l2": "inputs larger than L2 (no flush needed)"
                           if ncells * (lenA + lenb) * 8 > 256e6 else "inputs smaller than L2 (launch-bound configuration)",
                           "kernel": plan.kernel_name + (" + fused scatter into nzval" if fused else ", then gather_nzval"),
                           "nnz_per_gpu": int(slab.nnz) * nchunk},
                "clocks": clk.summary(), "e2e": e2e, "e2e_pageable": e2e_pageable, "e2e_affine_family": e2e_affine,
                "gpu_launches": int(launches),
                "roofline": roof, "cpu_baseline": cpu_base, "backsub": backsub, "device_generated_records": devgen,
                "multi_gpu_check": check, ("two_kernel_step" if fused else "fused_assembly"): other, "cuda_graph_step": graph_leg}
        print(json.dumps(line))
    if world > 1:
        ctx.comm_destroy()
        dist.destroy_process_group()


def main():
    # keep stdout clean for the ONE JSON line: libraries (NCCL's version banner, ...) write to fd 1 too
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w")
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C3", choices=sorted(CONFIGS))
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"])
    ap.add_argument("--dims", type=int, nargs="+", default=None)
    ap.add_argument("--e2e-dims", type=int, nargs="+", default=None)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--fused", type=int, default=0,
                    help="1: the condensation kernel scatters S_K into nzval itself (no gather pass); 0: condense, then gather. "
                         "The default line is the two-kernel step (its dominant kernel is the one the roofline describes); the "
                         "fused step is measured next to it and reported under fused_assembly")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
