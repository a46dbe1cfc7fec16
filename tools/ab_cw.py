"""A/B of the one-warp-per-cell condensation kernel (condense_cw.cu) against the 4-warps-per-cell left-looking kernel
(option cw = 0) : throughput with CUDA events on inputs larger than L2, difference of the results, keep_factors timing.
usage: [GHB_LIB_PATH=tools/_bin/libghb_x.so] python tools/ab_cw.py [ncells_log2=20] [shapes=34,36;33,12]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gridaphybrid_b200 as gh  # noqa: E402

RTH = np.array([[1, 1, 1], [1, 0, 0], [1, 0, 0]], bool)
ONES = np.ones((3, 3), bool)
SHAPES = {
    "34,36": ([30, 4, 36], ONES), "33,12": ([24, 9, 12], RTH), "40,36": ([30, 10, 36], ONES), "21,16": ([9, 12, 16], ONES),
    "56,16": ([40, 16, 16], RTH),
}
lg = int(sys.argv[1]) if len(sys.argv) > 1 else 20
which = sys.argv[2].split(";") if len(sys.argv) > 2 else ["34,36"]
ctx = gh.Context(0)
ev = lambda: torch.cuda.Event(enable_timing=True)


def timeit(plan, n, A, b, keep=False, reps=5):
    S = torch.empty((n, plan.n_b ** 2), dtype=torch.float64, device="cuda")
    g = torch.empty((n, plan.n_b), dtype=torch.float64, device="cuda")
    info = torch.empty(n, dtype=torch.int32, device="cuda")
    for _ in range(2):
        ctx.condense(plan, n, A, b, S, g, info, keep_factors=keep)
    e0, e1 = ev(), ev()
    e0.record()
    for _ in range(reps):
        ctx.condense(plan, n, A, b, S, g, info, keep_factors=keep)
    e1.record()
    torch.cuda.synchronize()
    return n / (e0.elapsed_time(e1) / reps) / 1e3, S, g, info


for name in which:
    if name in SHAPES:
        ndofs, touched = SHAPES[name]
        interior, boundary = [1, 2], [3]
    else:                                   # a named configuration of tests/helpers.py (any fields)
        from tests.helpers import CONFIGS
        c = CONFIGS[name]
        ndofs, touched, interior, boundary = c["ndofs"], c["touched"], c["interior"], c["boundary"]
    n = 1 << lg
    ctx.set_option("cw", 0)
    p_old = ctx.plan_blocks(ndofs, touched, interior, boundary)
    ctx.set_option("cw", 1)
    p_new = ctx.plan_blocks(ndofs, touched, interior, boundary)
    A = torch.empty((n, p_new.lenA), dtype=torch.float64, device="cuda")
    b = torch.empty((n, p_new.lenb), dtype=torch.float64, device="cuda")
    ctx.synth_fill(p_new, 0, n, A, b)
    r_new, S1, g1, i1 = timeit(p_new, n, A, b)
    r_old, S0, g0, i0 = timeit(p_old, n, A, b)
    dS = ((S1 - S0).norm(dim=1) / S0.norm(dim=1)).max().item()
    dg = ((g1 - g0).norm(dim=1) / g0.norm(dim=1)).max().item()
    print(f"({name}) {n} cells: new [{p_new.kernel_name}] {r_new:.2f} M cells/s, old [{p_old.kernel_name}] {r_old:.2f} M cells/s; "
          f"max rel diff S {dS:.2e} g {dg:.2e}; info new {int(i1.abs().sum())} old {int(i0.abs().sum())}", flush=True)
    del S0, g0, S1, g1
    try:
        rk, *_ = timeit(p_new, min(n, 1 << 19), A, b, keep=True)
        print(f"({name}) keep_factors new: {rk:.2f} M cells/s", flush=True)
    except Exception as e:  # noqa: BLE001
        print("keep_factors failed:", e)
