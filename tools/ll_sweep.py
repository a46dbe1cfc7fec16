"""Throughput of the left-looking (34,36) condensation kernel against the number of resident CTAs per SM."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gridaphybrid_b200 as gh  # noqa: E402

os.environ["GHB_DMMA_LL"] = os.environ.get("GHB_DMMA_LL", "1")
ctx = gh.Context(0)
ctx.set_option("cw", 0)      # these experiments are about the 4-warps-per-cell kernels
n = 1 << 19
plan = ctx.plan_blocks([30, 4, 36], np.ones((3, 3), bool), [1, 2], [3])
A = torch.empty((n, plan.lenA), dtype=torch.float64, device="cuda"); b = torch.empty((n, plan.lenb), dtype=torch.float64, device="cuda")
ctx.synth_fill(plan, 0, n, A, b)
S = torch.empty((n, plan.n_b ** 2), dtype=torch.float64, device="cuda"); g = torch.empty((n, plan.n_b), dtype=torch.float64, device="cuda")
info = torch.empty(n, dtype=torch.int32, device="cuda")
ev = lambda: torch.cuda.Event(enable_timing=True)


def timed(f, reps=3):
    f(); torch.cuda.synchronize()
    e0, e1 = ev(), ev(); e0.record()
    for _ in range(reps):
        f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


print("| CTAs/SM | condense M cells/s | cycles per cell latency (1.965 GHz) |")
print("|---|---|---|")
for k in range(1, 9):
    ctx.set_option("max_ctas_per_sm", k)
    mc = timed(lambda: ctx.condense(plan, n, A, b, S, g, info))
    print(f"| {k} | {n / mc / 1e3:.1f} | {mc * 1e-3 * 1.965e9 * 148 * k / n:.0f} |", flush=True)
