"""Driver for per-kernel ncu captures (profiles/r02_kernels.md): runs ONE kernel family a few times at a size larger than L2.
usage: python tools/prof_kernels.py <large|warp|gather|backsub_dmma|backsub_factors|batched_solve|cw33|expand> """
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import gridaphybrid_b200 as gh

which = sys.argv[1]
ctx = gh.Context(0)
RTH = np.array([[1, 1, 1], [1, 0, 0], [1, 0, 0]], bool)
ONES = np.ones((3, 3), bool)


def records(plan, n):
    A = torch.empty((n, plan.lenA), dtype=torch.float64, device="cuda")
    b = torch.empty((n, plan.lenb), dtype=torch.float64, device="cuda")
    ctx.synth_fill(plan, 0, n, A, b)
    return A, b


def condense(ndofs, touched, n, interior=(1, 2), boundary=(3,), reps=3):
    plan = ctx.plan_blocks(ndofs, touched, list(interior), list(boundary))
    A, b = records(plan, n)
    S = torch.empty((n, plan.n_b ** 2), dtype=torch.float64, device="cuda")
    g = torch.empty((n, plan.n_b), dtype=torch.float64, device="cuda")
    info = torch.empty(n, dtype=torch.int32, device="cuda")
    for _ in range(reps):
        ctx.condense(plan, n, A, b, S, g, info)
    torch.cuda.synchronize()
    print(which, plan.kernel_name, n, "cells, info", int(info.abs().sum()))
    return plan, A, b, S, g


if which == "large":
    condense([60, 60, 108], ONES, 8192)
elif which == "warp":
    condense([12, 4, 8], RTH, 1 << 21)
elif which == "warp78":
    condense([6, 1, 8], ONES, 1 << 22)
elif which == "cw33":
    condense([24, 9, 12], RTH, 1 << 19)
elif which == "gather":
    dims = (64, 64, 32)
    sk = gh.CartesianSkeleton(dims, ctx)
    M = gh.FacetFESpace(sk, 6, sk.facet_is_boundary())
    assem = gh.SparseMatrixAssembler(M)
    _, _, nnz = assem.symbolic()
    n = sk.ncells
    S = torch.randn((n, 36 * 36), dtype=torch.float64, device="cuda"); g = torch.randn((n, 36), dtype=torch.float64, device="cuda")
    nz = torch.empty(nnz, dtype=torch.float64, device="cuda"); rhs = torch.empty(assem.nrows, dtype=torch.float64, device="cuda")
    for _ in range(3):
        ctx.assemble_numeric(S, g, None, nz, rhs)
    torch.cuda.synchronize()
    print("gather", n, "cells", nnz, "nnz")
elif which in ("backsub_dmma", "backsub_factors"):
    ctx.set_option("cw", 0)
    plan = ctx.plan_blocks([30, 4, 36], ONES, [1, 2], [3])
    n = 1 << 17
    A, b = records(plan, n)
    ids = torch.randint(1, 5000, (n, 36), dtype=torch.int64, device="cuda")
    lam = torch.randn(5000, dtype=torch.float64, device="cuda")
    u = torch.empty((n, 34), dtype=torch.float64, device="cuda")
    if which == "backsub_factors":
        S = torch.empty((n, 36 * 36), dtype=torch.float64, device="cuda"); g = torch.empty((n, 36), dtype=torch.float64, device="cuda")
        ctx.condense(plan, n, A, b, S, g, None, keep_factors=True)
        for _ in range(3):
            ctx.backsub(plan, n, None, None, lam, None, ids, u, None)
    else:
        for _ in range(3):
            ctx.backsub(plan, n, A, b, lam, None, ids, u, None)
    torch.cuda.synchronize()
    print(which, plan.kernel_name, n)
elif which == "batched_solve":
    nb, n, m = 1 << 20, 6, 30
    A = torch.randn((nb, n * n), dtype=torch.float64, device="cuda") + 4 * torch.eye(n, dtype=torch.float64, device="cuda").reshape(1, -1)
    B = torch.randn((nb, n * m), dtype=torch.float64, device="cuda"); X = torch.empty_like(B)
    for _ in range(3):
        ctx.l2_projection_dofs(nb, n, m, A, B, X, None)
    torch.cuda.synchronize()
    print("batched_solve", nb)
elif which == "expand":
    plan = ctx.plan_blocks([30, 4, 36], ONES, [1, 2], [3])
    n, ntab = 1 << 17, 7
    TA = torch.randn((ntab, plan.lenA), dtype=torch.float64, device="cuda"); Tb = torch.randn((ntab, plan.lenb), dtype=torch.float64, device="cuda")
    coef = torch.randn((n, ntab), dtype=torch.float64, device="cuda")
    A = torch.empty((n, plan.lenA), dtype=torch.float64, device="cuda"); b = torch.empty((n, plan.lenb), dtype=torch.float64, device="cuda")
    for _ in range(3):
        ctx.expand_records(plan, n, ntab, TA, Tb, coef, A, b)
    torch.cuda.synchronize()
    print("expand", n)
elif which in ("cw_back", "cw_gen", "cw_gen_back"):
    plan = ctx.plan_blocks([30, 4, 36], ONES, [1, 2], [3])
    n, ntab = 1 << 17, 7
    A, b = records(plan, n)
    rng = np.random.default_rng(0)
    fam = gh.AffineRecordFamily(np.concatenate([A[:1].cpu().numpy(), 1e-3 * rng.standard_normal((ntab - 1, plan.lenA))]),
                                np.concatenate([b[:1].cpu().numpy(), 1e-3 * rng.standard_normal((ntab - 1, plan.lenb))]))
    coef = torch.cat([torch.ones((n, 1), dtype=torch.float64, device="cuda"), torch.rand((n, ntab - 1), dtype=torch.float64, device="cuda")], dim=1)
    ids = torch.randint(1, 100001, (n, plan.n_b), device="cuda", dtype=torch.int64)
    lam = torch.randn(100000, dtype=torch.float64, device="cuda")
    u = torch.empty((n, plan.n_i), dtype=torch.float64, device="cuda")
    S = torch.empty((n, plan.n_b ** 2), dtype=torch.float64, device="cuda"); g = torch.empty((n, plan.n_b), dtype=torch.float64, device="cuda")
    info = torch.empty(n, dtype=torch.int32, device="cuda")
    for _ in range(3):
        if which == "cw_back":
            ctx.backsub(plan, n, A, b, lam, None, ids, u, info)
        elif which == "cw_gen":
            fam.condense(ctx, plan, coef, S, g, info)
        else:
            fam.backsub(ctx, plan, coef, lam, None, ids, u, info)
    torch.cuda.synchronize()
    print(which, n, "info", int(info.abs().sum()))
