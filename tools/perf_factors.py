"""Forward condensation with keep_factors and the backward map from the stored factors on the DMMA
shapes, next to the plain forward kernel and the backward kernel that recomputes the LU (CUDA events)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gridaphybrid_b200 as gh  # noqa: E402

RTH = np.array([[1, 1, 1], [1, 0, 0], [1, 0, 0]], bool)
SHAPES = {"(34,36)": ([30, 4, 36], np.ones((3, 3), bool), 1 << 19), "(33,12)": ([24, 9, 12], RTH, 1 << 19),
          "(56,16)": ([40, 16, 16], RTH, 1 << 18)}
ctx = gh.Context(0)
ev = lambda: torch.cuda.Event(enable_timing=True)


def timed(f, reps=3):
    f(); torch.cuda.synchronize()
    e0, e1 = ev(), ev(); e0.record()
    for _ in range(reps):
        f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


print("| shape | forward | forward keep_factors | forward keep_factors (generic kernel) | backward (LU recomputed) | backward from factors | (M cells/s) |")
print("|---|---|---|---|---|---|---|")
for name, (ndofs, touched, n) in SHAPES.items():
    plan = ctx.plan_blocks(ndofs, touched, [1, 2], [3])
    A = torch.empty((n, plan.lenA), dtype=torch.float64, device="cuda"); b = torch.empty((n, plan.lenb), dtype=torch.float64, device="cuda")
    ctx.synth_fill(plan, 0, n, A, b)
    S = torch.empty((n, plan.n_b ** 2), dtype=torch.float64, device="cuda"); g = torch.empty((n, plan.n_b), dtype=torch.float64, device="cuda")
    lam = torch.randn(n * plan.n_b, dtype=torch.float64, device="cuda")
    ids = torch.arange(1, n * plan.n_b + 1, dtype=torch.int64, device="cuda")
    u = torch.empty((n, plan.n_i), dtype=torch.float64, device="cuda")
    f0 = timed(lambda: ctx.condense(plan, n, A, b, S, g, None))
    ctx.set_option("factors_generic", 1)
    f2 = timed(lambda: ctx.condense(plan, n, A, b, S, g, None, keep_factors=True), reps=1)
    ctx.set_option("factors_generic", 0)
    f1 = timed(lambda: ctx.condense(plan, n, A, b, S, g, None, keep_factors=True))
    b1 = timed(lambda: ctx.backsub(plan, n, None, None, lam, None, ids, u, None))
    b0 = timed(lambda: ctx.backsub(plan, n, A, b, lam, None, ids, u, None))
    r = lambda ms: f"{n / ms / 1e3:.1f}"
    print(f"| {name} | {r(f0)} | {r(f1)} | {r(f2)} | {r(b0)} | {r(b1)} | |", flush=True)
    del A, b, S, g, lam, ids, u
    torch.cuda.empty_cache()
