"""Backward map (LU recomputed per cell, as the reference does): the cell-warp BACK kernel against the plan's other
backward kernel (option cw_back = 0: 4-warps-per-cell DMMA or generic), M cells/s per shape."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gridaphybrid_b200 as gh  # noqa: E402
from tests.helpers import CONFIGS  # noqa: E402

ctx = gh.Context(0)
names = sys.argv[1:] or ["C3_hdg_k2_3d", "C2_rth_k2_2d", "hdg_equal_order_3d", "elasticity_k1_2d", "hencky_k1_2d", "rth_k0_2d"]
ev = lambda: torch.cuda.Event(enable_timing=True)
print("| shape | (n_i, n_b) | kernel | cells | BACK M cells/s | other M cells/s |")
print("|---|---|---|---|---|---|")
for name in names:
    c = CONFIGS[name]
    plan = ctx.plan_blocks(c["ndofs"], c["touched"], c["interior"], c["boundary"])
    n = int(min(2 ** 20, 12e9 // ((plan.lenA + plan.lenb) * 8)))
    A = torch.empty((n, plan.lenA), dtype=torch.float64, device="cuda"); b = torch.empty((n, plan.lenb), dtype=torch.float64, device="cuda")
    ctx.synth_fill(plan, 0, n, A, b)
    nfree = 100000
    ids = torch.randint(1, nfree + 1, (n, plan.n_b), device="cuda", dtype=torch.int64)
    lam = torch.randn(nfree, dtype=torch.float64, device="cuda")
    u = torch.empty((n, plan.n_i), dtype=torch.float64, device="cuda")
    info = torch.empty(n, dtype=torch.int32, device="cuda")
    rates = []
    for cw_back in (1, 0):
        ctx.set_option("cw_back", cw_back)
        ctx.backsub(plan, n, A, b, lam, None, ids, u, info); torch.cuda.synchronize()
        e0, e1 = ev(), ev(); e0.record()
        for _ in range(3):
            ctx.backsub(plan, n, A, b, lam, None, ids, u, info)
        e1.record(); torch.cuda.synchronize()
        rates.append(n / (e0.elapsed_time(e1) / 3) / 1e3)
        assert int(info.abs().sum()) == 0
    print(f"| {name} | ({plan.n_i}, {plan.n_b}) | {plan.kernel_name} | {n} | {rates[0]:.1f} | {rates[1]:.1f} |")
    del A, b, ids, u
