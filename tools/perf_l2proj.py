"""Throughput of ghb_l2_projection_dofs_f64 (batched A\\B) on the facet sizes of the named configurations."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gridaphybrid_b200 as gh  # noqa: E402

ctx = gh.Context(0)
ev = lambda: torch.cuda.Event(enable_timing=True)
print("| n (facet dofs) | nrhs (bulk dofs) | systems | M systems/s | GB/s (8(n^2 + 2 n nrhs) per system) |")
print("|---|---|---|---|---|")
for n, m, nb in [(6, 30, 1 << 21), (18, 60, 1 << 19), (18, 1, 1 << 21), (9, 30, 1 << 20)]:
    A = torch.randn((nb, n * n), dtype=torch.float64, device="cuda")
    A.view(nb, n, n).diagonal(dim1=1, dim2=2).add_(2.0 * n ** 0.5)
    B = torch.randn((nb, n * m), dtype=torch.float64, device="cuda")
    X = torch.empty_like(B)
    for _ in range(2):
        ctx.l2_projection_dofs(nb, n, m, A, B, X, None)
    e0, e1 = ev(), ev(); e0.record()
    for _ in range(3):
        ctx.l2_projection_dofs(nb, n, m, A, B, X, None)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"| {n} | {m} | {nb} | {nb / ms / 1e3:.1f} | {8 * (n * n + 2 * n * m) * nb / ms / 1e6:.0f} |", flush=True)
    del A, B, X
