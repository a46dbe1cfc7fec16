"""Small driver for compute-sanitizer: every tuned kernel once on a few cells (memcheck / racecheck / synccheck)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gridaphybrid_b200 as gh  # noqa: E402

ctx = gh.Context(0)
shapes = {"(7,8)": ([6, 1, 8], np.ones((3, 3), bool)), "(16,8)": ([12, 4, 8], np.array([[1, 1, 1], [1, 0, 0], [1, 0, 0]], bool)),
          "(33,12)": ([24, 9, 12], np.array([[1, 1, 1], [1, 0, 0], [1, 0, 0]], bool)),
          "(56,16)": ([40, 16, 16], np.array([[1, 1, 1], [1, 0, 0], [1, 0, 0]], bool)),
          "(34,36)": ([30, 4, 36], np.ones((3, 3), bool)), "(5+7,3) generic": ([5, 3, 7], np.ones((3, 3), bool))}
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
for name, (ndofs, touched) in shapes.items():
    plan = ctx.plan_blocks(ndofs, touched, [1, 2] if "generic" not in name else [3, 1], [3] if "generic" not in name else [2])
    A = torch.empty((n, plan.lenA), dtype=torch.float64, device="cuda"); b = torch.empty((n, plan.lenb), dtype=torch.float64, device="cuda")
    ctx.synth_fill(plan, 0, n, A, b)
    S = torch.empty((n, plan.n_b ** 2), dtype=torch.float64, device="cuda"); g = torch.empty((n, plan.n_b), dtype=torch.float64, device="cuda")
    info = torch.empty(n, dtype=torch.int32, device="cuda")
    ctx.condense(plan, n, A, b, S, g, info)
    lam = torch.randn(n * plan.n_b, dtype=torch.float64, device="cuda")
    ids = torch.arange(1, n * plan.n_b + 1, dtype=torch.int64, device="cuda")
    u = torch.empty((n, plan.n_i), dtype=torch.float64, device="cuda")
    ctx.backsub(plan, n, A, b, lam, None, ids, u, info)
    ctx.condense(plan, n, A, b, S, g, info, keep_factors=True)      # DMMA shapes: left-looking kernel + back substitution
    ctx.backsub(plan, n, None, None, lam, None, ids, u, None)        # backward map from the stored factors
    if plan.kernel_name.startswith("dmma"):
        ctx.set_option("dmma_ll", 0)                              # the right-looking kernel too
        ctx.condense(plan, n, A, b, S, g, info)
        ctx.set_option("dmma_ll", 1)
    torch.cuda.synchronize()
    print(name, plan.kernel_name, "ok", float(S.abs().max()), float(u.abs().max()))
# assembly on a small mesh
sk = gh.CartesianSkeleton((6, 5, 4), ctx)
M = gh.FacetFESpace(sk, 6, sk.facet_is_boundary())
asm = gh.SparseMatrixAssembler(M)
colptr, rowval, nnz = asm.symbolic()
nc = asm.cell_ids.shape[0]
S = torch.randn((nc, 36 * 36), dtype=torch.float64, device="cuda"); g = torch.randn((nc, 36), dtype=torch.float64, device="cuda")
nz = torch.empty(nnz, dtype=torch.float64, device="cuda"); rhs = torch.empty(asm.nrows, dtype=torch.float64, device="cuda")
ctx.assemble_numeric(S, g, M.dirichlet_values, nz, rhs)
ctx.assemble_numeric_csr(S, g, M.dirichlet_values, nz, rhs)        # CSR hand-off: in-place block transpose + gather
torch.cuda.synchronize()
print("assembly ok", nnz)
# SURVEY 8f kernels: batched A\B (factors in shared memory, column per lane), affine-family record expansion
for n_, m_ in [(6, 30), (18, 60), (18, 1), (32, 9)]:
    nb = 300
    Aq = torch.randn((nb, n_ * n_), dtype=torch.float64, device="cuda")
    Aq.view(nb, n_, n_).diagonal(dim1=1, dim2=2).add_(2.0 * n_ ** 0.5)
    Bq = torch.randn((nb, n_ * m_), dtype=torch.float64, device="cuda")
    Xq = torch.empty_like(Bq); iq = torch.empty(nb, dtype=torch.int32, device="cuda")
    ctx.l2_projection_dofs(nb, n_, m_, Aq, Bq, Xq, iq)
    torch.cuda.synchronize()
    print("l2 projection", n_, m_, "ok", int(iq.abs().sum()))
plan = ctx.plan_blocks([30, 4, 36], np.ones((3, 3), bool), [1, 2], [3])
fam = gh.AffineRecordFamily(np.random.default_rng(0).standard_normal((7, plan.lenA)), np.random.default_rng(1).standard_normal((7, plan.lenb)))
cells = fam.expand(ctx, plan, gh.cartesian_coefficients((7, 6, 5), (0.1, 0.1, 0.1), "cuda"))
torch.cuda.synchronize()
print("expand ok", float(cells.A.abs().max()))
