"""Condensation with the records of an affine family generated in the loader (ghb_condense_affine_f64)
against expand -> condense with records in HBM: bit-equality and throughput, C3 shape by default."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gridaphybrid_b200 as gh  # noqa: E402

ctx = gh.Context(0)
if os.environ.get("MAXCTAS"):
    ctx.set_option("max_ctas_per_sm", int(os.environ["MAXCTAS"]))
dims = tuple(int(x) for x in os.environ.get("DIMS", "96,96,96").split(","))
n = int(np.prod(dims))
shape = os.environ.get("SHAPE", "C3")
if shape == "C3":
    ndofs, touched, I, B = [30, 4, 36], np.ones((3, 3), bool), [1, 2], [3]
elif shape == "RTH2":     # (33,12): RT-H k=2 2-D, untouched (p,p) and (lambda, p)/(p, lambda) blocks
    t = np.ones((3, 3), bool); t[1, 1] = False; t[1, 2] = False; t[2, 1] = False
    ndofs, touched, I, B = [24, 9, 12], t, [1, 2], [3]
if shape not in ("C3", "RTH2"):
    if shape == "HDG3D_K1":       # Darcy HDG k=1 on hexes: u 3 x 8, p 8 | lambda 6 x 4 -> (32, 24), shape-generic class 32
        ndofs, touched, I, B = [24, 8, 24], np.ones((3, 3), bool), [1, 2], [3]
    else:
        from tests.helpers import CONFIGS
        c = CONFIGS[shape]
        ndofs, touched, I, B = c["ndofs"], c["touched"], c["interior"], c["boundary"]
plan = ctx.plan_blocks(ndofs, touched, I, B)
ntab = int(os.environ.get("NTAB", "7"))
rng = np.random.default_rng(0)
A0 = torch.empty((1, plan.lenA), dtype=torch.float64, device="cuda"); b0 = torch.empty((1, plan.lenb), dtype=torch.float64, device="cuda")
ctx.synth_fill(plan, 0, 1, A0, b0)
TA = np.concatenate([A0.cpu().numpy(), 1e-3 * rng.standard_normal((ntab - 1, plan.lenA))])
Tb = np.concatenate([b0.cpu().numpy(), 1e-3 * rng.standard_normal((ntab - 1, plan.lenb))])
fam = gh.AffineRecordFamily(TA, Tb)
coef = torch.cat([torch.ones((n, 1), dtype=torch.float64, device="cuda"),
                  torch.rand((n, ntab - 1), dtype=torch.float64, device="cuda")], dim=1).contiguous()
A = torch.empty((n, plan.lenA), dtype=torch.float64, device="cuda"); b = torch.empty((n, plan.lenb), dtype=torch.float64, device="cuda")
S = torch.empty((n, plan.n_b ** 2), dtype=torch.float64, device="cuda"); g = torch.empty((n, plan.n_b), dtype=torch.float64, device="cuda")
S2 = torch.empty_like(S); g2 = torch.empty_like(g)
info = torch.empty(n, dtype=torch.int32, device="cuda"); info2 = torch.empty_like(info)
ev = lambda: torch.cuda.Event(enable_timing=True)


def timed(f, reps=3):
    f(); torch.cuda.synchronize()
    e0, e1 = ev(), ev(); e0.record()
    for _ in range(reps):
        f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def two_step():
    fam.expand(ctx, plan, coef, A, b)
    ctx.condense(plan, n, A, b, S, g, info)


def fused():
    fam.condense(ctx, plan, coef, S2, g2, info2)


print(f"kernel {plan.kernel_name}, {n} cells, {ntab} tables")
if os.environ.get("ONLY") == "fused":          # ncu captures: nothing but the GEN kernel
    for _ in range(3):
        fused()
    torch.cuda.synchronize()
    sys.exit(0)
tc = timed(lambda: ctx.condense(plan, n, A, b, S, g, info)) if True else 0
t2 = timed(two_step)
tf = timed(fused)
eq = bool(torch.equal(S, S2) and torch.equal(g, g2) and torch.equal(info, info2))
print(f"condense (resident records):     {tc:.2f} ms = {n / tc / 1e3:.1f} M cells/s")
print(f"expand + condense (records in HBM): {t2:.2f} ms = {n / t2 / 1e3:.1f} M cells/s")
print(f"condense_affine (records in the loader): {tf:.2f} ms = {n / tf / 1e3:.1f} M cells/s   bit-equal: {eq}  bad info: {int(info2.abs().sum())}")
if not eq:
    d = (S - S2).abs().max().item()
    print("max |dS|", d, "nan:", int(torch.isnan(S2).sum()))
    sys.exit(1)
# backward map from coefficients
nfree = 100000
ids = torch.randint(1, nfree + 1, (n, plan.n_b), device="cuda", dtype=torch.int64)
lam = torch.randn(nfree, dtype=torch.float64, device="cuda")
u = torch.empty((n, plan.n_i), dtype=torch.float64, device="cuda"); u2 = torch.empty_like(u)
tb0 = timed(lambda: ctx.backsub(plan, n, A, b, lam, None, ids, u, info))


def back2():
    fam.expand(ctx, plan, coef, A, b)
    ctx.backsub(plan, n, A, b, lam, None, ids, u, info)


tb2 = timed(back2)
tbf = timed(lambda: fam.backsub(ctx, plan, coef, lam, None, ids, u2, info2))
print(f"backsub (resident records):   {tb0:.2f} ms = {n / tb0 / 1e3:.1f} M cells/s")
print(f"expand + backsub:             {tb2:.2f} ms = {n / tb2 / 1e3:.1f} M cells/s")
print(f"backsub_affine (in the loader): {tbf:.2f} ms = {n / tbf / 1e3:.1f} M cells/s   bit-equal: {bool(torch.equal(u, u2))}")
if os.environ.get("ASSEMBLE", "1") == "1" and shape == "C3":
    sk = gh.CartesianSkeleton(dims, ctx)
    M = gh.FacetFESpace(sk, 6, sk.facet_is_boundary())
    asm = gh.SparseMatrixAssembler(M)
    colptr, rowval, nnz = asm.symbolic()
    nz = torch.empty(nnz, dtype=torch.float64, device="cuda"); rhs = torch.empty(asm.nrows, dtype=torch.float64, device="cuda")
    nz2 = torch.empty_like(nz); rhs2 = torch.empty_like(rhs)

    def pipe2():
        fam.expand(ctx, plan, coef, A, b)
        ctx.condense_assemble(plan, n, A, b, None, nz, rhs, info)

    ta = timed(lambda: ctx.condense_assemble(plan, n, A, b, None, nz, rhs, info))
    tp2 = timed(pipe2)
    tpf = timed(lambda: fam.condense_assemble(ctx, plan, coef, None, nz2, rhs2, info2))
    eq2 = bool(torch.equal(nz, nz2) and torch.equal(rhs, rhs2))
    print(f"condense_assemble (resident records): {ta:.2f} ms = {n / ta / 1e3:.1f} M cells/s")
    print(f"expand + condense_assemble:           {tp2:.2f} ms = {n / tp2 / 1e3:.1f} M cells/s")
    print(f"condense_assemble_affine:             {tpf:.2f} ms = {n / tpf / 1e3:.1f} M cells/s   bit-equal: {eq2}")
    if not eq2:
        sys.exit(1)
