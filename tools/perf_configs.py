"""Condensation / back-substitution throughput of the named BASELINE.json shapes on synthetic records
(CUDA events, inputs larger than L2), with the HBM roofline of each shape."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gridaphybrid_b200 as gh  # noqa: E402

CONFIGS = {
    "C1 Darcy HDG k=1 2-D (7,8)": ([6, 1, 8], np.ones((3, 3), bool), 4 << 20),
    "C2 Darcy RT-H k=1 2-D (16,8)": ([12, 4, 8], np.array([[1, 1, 1], [1, 0, 0], [1, 0, 0]], bool), 4 << 20),
    "C2 Darcy RT-H k=2 2-D (33,12)": ([24, 9, 12], np.array([[1, 1, 1], [1, 0, 0], [1, 0, 0]], bool), 1 << 20),
    "C2 Darcy RT-H k=3 2-D (56,16)": ([40, 16, 16], np.array([[1, 1, 1], [1, 0, 0], [1, 0, 0]], bool), 1 << 19),
    "C3 Darcy HDG k=2 3-D (34,36)": ([30, 4, 36], np.ones((3, 3), bool), 1 << 20),
    "C4 Elasticity HDG k=2 3-D (120,108)": ([60, 60, 108], np.ones((3, 3), bool), 1 << 15),
    "C5 Hencky HDG k=1 3-D (106,72)": ([12, 12, 4, 24, 24, 30, 18, 54], np.ones((8, 8), bool), 1 << 16),
}
FIELDS = {"C5 Hencky HDG k=1 3-D (106,72)": ([1, 2, 3, 4, 5, 6], [7, 8])}
if os.environ.get("GHB_PERF_ONLY"):
    CONFIGS = {k: v for k, v in CONFIGS.items() if os.environ["GHB_PERF_ONLY"] in k}
peak = 6545.9
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass

ctx = gh.Context(0)
ev = lambda: torch.cuda.Event(enable_timing=True)
print(f"| config | kernel | cells | condense M cells/s | of roofline (slower of HBM, FP64) | backsub M cells/s |")
print("|---|---|---|---|---|---|")
for name, (ndofs, touched, n) in CONFIGS.items():
    plan = ctx.plan_blocks(ndofs, touched, *FIELDS.get(name, ([1, 2], [3])))
    A = torch.empty((n, plan.lenA), dtype=torch.float64, device="cuda"); b = torch.empty((n, plan.lenb), dtype=torch.float64, device="cuda")
    ctx.synth_fill(plan, 0, n, A, b)
    S = torch.empty((n, plan.n_b ** 2), dtype=torch.float64, device="cuda"); g = torch.empty((n, plan.n_b), dtype=torch.float64, device="cuda")
    info = torch.empty(n, dtype=torch.int32, device="cuda")
    for _ in range(3):
        ctx.condense(plan, n, A, b, S, g, info)
    e0, e1 = ev(), ev(); e0.record()
    for _ in range(5):
        ctx.condense(plan, n, A, b, S, g, info)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    bytes_cell = 8 * (plan.lenA + plan.lenb + plan.n_b ** 2 + plan.n_b)
    ni, nb_ = plan.n_i, plan.n_b
    flops_cell = 2 * ni ** 3 / 3 + 2 * ni ** 2 * (nb_ + 1) + 2 * ni * nb_ * (nb_ + 1)
    roof = min(peak * 1e9 / bytes_cell, 37.1e12 / flops_cell)     # slower of HBM and FP64 (DMMA, measured)
    lam = torch.randn(n * plan.n_b, dtype=torch.float64, device="cuda")
    ids = torch.arange(1, n * plan.n_b + 1, dtype=torch.int64, device="cuda")
    u = torch.empty((n, plan.n_i), dtype=torch.float64, device="cuda")
    for _ in range(2):
        ctx.backsub(plan, n, A, b, lam, None, ids, u, info)
    e0, e1 = ev(), ev(); e0.record()
    for _ in range(3):
        ctx.backsub(plan, n, A, b, lam, None, ids, u, info)
    e1.record(); torch.cuda.synchronize()
    msb = e0.elapsed_time(e1) / 3
    print(f"| {name} | {plan.kernel_name} | {n} | {n / ms / 1e3:.2f} | {n / ms * 1e3 / roof:.3f} | {n / msb / 1e3:.2f} |")
    assert int(info.abs().sum()) == 0
    del A, b, S, g, lam, ids, u
    torch.cuda.empty_cache()
