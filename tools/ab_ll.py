"""A/B of the two DMMA condensation kernels (right-looking bottom block in shared memory vs left-looking bottom block in
registers, GHB_DMMA_LL) on the three DMMA shapes: throughput with CUDA events on inputs larger than L2, and the
difference of their results (both are checked against the oracle by tests/test_gpu_parity.py)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gridaphybrid_b200 as gh  # noqa: E402

RTH = np.array([[1, 1, 1], [1, 0, 0], [1, 0, 0]], bool)
SHAPES = {
    "(34,36)": ([30, 4, 36], np.ones((3, 3), bool), 1 << 20),
    "(33,12)": ([24, 9, 12], RTH, 1 << 20),
    "(56,16)": ([40, 16, 16], RTH, 1 << 19),
}
only = os.environ.get("GHB_AB_ONLY")
ctx = gh.Context(0)
ev = lambda: torch.cuda.Event(enable_timing=True)


def run(plan, n, A, b, env):
    for k in ("GHB_DMMA_LL", "GHB_LL_CTAS"):
        os.environ.pop(k, None)
    os.environ.update(env)
    S = torch.empty((n, plan.n_b ** 2), dtype=torch.float64, device="cuda")
    g = torch.empty((n, plan.n_b), dtype=torch.float64, device="cuda")
    info = torch.empty(n, dtype=torch.int32, device="cuda")
    for _ in range(3):
        ctx.condense(plan, n, A, b, S, g, info)
    e0, e1 = ev(), ev()
    e0.record()
    for _ in range(5):
        ctx.condense(plan, n, A, b, S, g, info)
    e1.record()
    torch.cuda.synchronize()
    assert int(info.abs().sum()) == 0
    return n / (e0.elapsed_time(e1) / 5) / 1e3, S, g


print("| shape | variant | M cells/s | max rel diff vs right-looking kernel |")
print("|---|---|---|---|")
for name, (ndofs, touched, n) in SHAPES.items():
    if only and only not in name:
        continue
    plan = ctx.plan_blocks(ndofs, touched, [1, 2], [3])
    A = torch.empty((n, plan.lenA), dtype=torch.float64, device="cuda")
    b = torch.empty((n, plan.lenb), dtype=torch.float64, device="cuda")
    ctx.synth_fill(plan, 0, n, A, b)
    r0, S0, g0 = run(plan, n, A, b, {"GHB_DMMA_LL": "0"})
    print(f"| {name} | right-looking (smem bottom block) | {r0:.2f} | |", flush=True)
    variants = [{"GHB_DMMA_LL": "1"}]
    if name == "(34,36)":
        variants = [{"GHB_DMMA_LL": "1", "GHB_LL_CTAS": str(c)} for c in (8, 7, 6)]
    for env in variants:
        r1, S1, g1 = run(plan, n, A, b, env)
        sc = S0.abs().amax(dim=1, keepdim=True)
        dS = float(((S1 - S0).abs() / sc).max())
        dg = float(((g1 - g0).abs() / g0.abs().amax(dim=1, keepdim=True)).max())
        print(f"| {name} | left-looking {env} | {r1:.2f} | S {dS:.2e}, g {dg:.2e} |", flush=True)
        del S1, g1
    del A, b, S0, g0
    torch.cuda.empty_cache()
