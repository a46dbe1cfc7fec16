"""Small driver for compute-sanitizer, round-2 kernels: the cell-warp kernel in every mode (condensation, keep_factors,
fused scatter, BACK, GEN, GEN + scatter, GEN + BACK; tuned shapes and shape-generic classes, persistent loop wrapping) and
the one-CTA-per-system batched solve."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gridaphybrid_b200 as gh  # noqa: E402
from tests.helpers import CONFIGS  # noqa: E402

ctx = gh.Context(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2600
rng = np.random.default_rng(0)
names = os.environ.get("SHAPES", "C3_hdg_k2_3d,C2_rth_k2_2d,hdg_equal_order_3d,elasticity_k1_2d,hencky_k1_2d,rth_k0_2d").split(",")
for name in names:
    c = CONFIGS[name]
    plan = ctx.plan_blocks(c["ndofs"], c["touched"], c["interior"], c["boundary"])
    A = torch.empty((n, plan.lenA), dtype=torch.float64, device="cuda"); b = torch.empty((n, plan.lenb), dtype=torch.float64, device="cuda")
    ctx.synth_fill(plan, 0, n, A, b)
    S = torch.empty((n, plan.n_b ** 2), dtype=torch.float64, device="cuda"); g = torch.empty((n, plan.n_b), dtype=torch.float64, device="cuda")
    info = torch.empty(n, dtype=torch.int32, device="cuda")
    ctx.condense(plan, n, A, b, S, g, info)
    ctx.condense(plan, n, A, b, S, g, info, keep_factors=True)
    ids = torch.randint(1, 5000, (n, plan.n_b), device="cuda", dtype=torch.int64)
    lam = torch.randn(5000, dtype=torch.float64, device="cuda")
    u = torch.empty((n, plan.n_i), dtype=torch.float64, device="cuda")
    ctx.backsub(plan, n, A, b, lam, None, ids, u, info)                       # BACK
    ntab = 7
    TA = np.concatenate([A[:1].cpu().numpy(), 1e-2 * rng.standard_normal((ntab - 1, plan.lenA))])
    Tb = np.concatenate([b[:1].cpu().numpy(), rng.standard_normal((ntab - 1, plan.lenb))])
    fam = gh.AffineRecordFamily(TA, Tb)
    coef = torch.cat([torch.ones((n, 1), dtype=torch.float64, device="cuda"), 0.1 * torch.rand((n, ntab - 1), dtype=torch.float64, device="cuda")], dim=1)
    fam.condense(ctx, plan, coef, S, g, info)                                # GEN (or the chunked fallback)
    fam.condense(ctx, plan, coef, S, g, info, keep_factors=True)             # GEN + KEEPX
    ctx.backsub(plan, n, None, None, lam, None, ids, u, None)                # backward map from those factors
    fam.backsub(ctx, plan, coef, lam, None, ids, u, info)                    # GEN + BACK
    print(name, plan.kernel_name, "info", int(info.abs().sum()))
# fused scatter, resident records and GEN, on a small mesh
plan = ctx.plan_blocks([30, 4, 36], np.ones((3, 3), bool), [1, 2], [3])
sk = gh.CartesianSkeleton((14, 14, 13), ctx)
M = gh.FacetFESpace(sk, 6, sk.facet_is_boundary())
asm = gh.SparseMatrixAssembler(M)
colptr, rowval, nnz = asm.symbolic()
m = sk.ncells
A = torch.empty((m, plan.lenA), dtype=torch.float64, device="cuda"); b = torch.empty((m, plan.lenb), dtype=torch.float64, device="cuda")
ctx.synth_fill(plan, 0, m, A, b)
nz = torch.empty(nnz, dtype=torch.float64, device="cuda"); rhs = torch.empty(asm.nrows, dtype=torch.float64, device="cuda")
info = torch.empty(m, dtype=torch.int32, device="cuda")
ctx.condense_assemble(plan, m, A, b, None, nz, rhs, info)                     # SCAT
fam = gh.AffineRecordFamily(np.concatenate([A[:1].cpu().numpy(), 1e-2 * rng.standard_normal((2, plan.lenA))]),
                            np.concatenate([b[:1].cpu().numpy(), rng.standard_normal((2, plan.lenb))]))
coef = torch.cat([torch.ones((m, 1), dtype=torch.float64, device="cuda"), 0.1 * torch.rand((m, 2), dtype=torch.float64, device="cuda")], dim=1)
fam.condense_assemble(ctx, plan, coef, None, nz, rhs, info)                   # GEN + SCAT
print("fused", m, "cells, info", int(info.abs().sum()))
# batched solve, n > 32
Q = rng.standard_normal((50, 45, 45)); Am = Q @ np.transpose(Q, (0, 2, 1)) / 45 + 0.5 * np.eye(45)
X = gh.compute_bulk_to_skeleton_l2_projection_dofs(Am, rng.standard_normal((50, 45, 40)), ctx)
torch.cuda.synchronize()
print("done")
