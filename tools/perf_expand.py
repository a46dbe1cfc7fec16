"""Throughput of ghb_expand_records_f64 (records of an affine family generated on the device) on the C3
record shape, and of the whole device pipeline expand -> condense -> assemble from cell coefficients."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gridaphybrid_b200 as gh  # noqa: E402

ctx = gh.Context(0)
dims = (96, 96, 96)
n = int(np.prod(dims))
plan = ctx.plan_blocks([30, 4, 36], np.ones((3, 3), bool), [1, 2], [3])
ntab = 7
rng = np.random.default_rng(0)
# synthetic tables: a well-conditioned base record plus small perturbation tables (values do not matter for the timing)
A0 = torch.empty((1, plan.lenA), dtype=torch.float64, device="cuda"); b0 = torch.empty((1, plan.lenb), dtype=torch.float64, device="cuda")
ctx.synth_fill(plan, 0, 1, A0, b0)
TA = np.concatenate([A0.cpu().numpy(), 1e-3 * rng.standard_normal((ntab - 1, plan.lenA))])
Tb = np.concatenate([b0.cpu().numpy(), 1e-3 * rng.standard_normal((ntab - 1, plan.lenb))])
fam = gh.AffineRecordFamily(TA, Tb)
coef = gh.cartesian_coefficients(dims, (1 / 96,) * 3, "cuda")
A = torch.empty((n, plan.lenA), dtype=torch.float64, device="cuda"); b = torch.empty((n, plan.lenb), dtype=torch.float64, device="cuda")
S = torch.empty((n, plan.n_b ** 2), dtype=torch.float64, device="cuda"); g = torch.empty((n, plan.n_b), dtype=torch.float64, device="cuda")
info = torch.empty(n, dtype=torch.int32, device="cuda")
sk = gh.CartesianSkeleton(dims, ctx)
M = gh.FacetFESpace(sk, 6, sk.facet_is_boundary())
asm = gh.SparseMatrixAssembler(M)
colptr, rowval, nnz = asm.symbolic()
nz = torch.empty(nnz, dtype=torch.float64, device="cuda"); rhs = torch.empty(asm.nrows, dtype=torch.float64, device="cuda")
ev = lambda: torch.cuda.Event(enable_timing=True)


def timed(f, reps=3):
    f(); torch.cuda.synchronize()
    e0, e1 = ev(), ev(); e0.record()
    for _ in range(reps):
        f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def pipeline():
    fam.expand(ctx, plan, coef, A, b)
    ctx.condense(plan, n, A, b, S, g, info)
    ctx.assemble_numeric(S, g, None, nz, rhs)


te = timed(lambda: fam.expand(ctx, plan, coef, A, b))
tp = timed(pipeline)
bytes_cell = 8 * (plan.lenA + plan.lenb)
print(f"expand: {n} cells, {ntab} tables: {te:.2f} ms = {n / te / 1e3:.1f} M cells/s = {bytes_cell * n / te / 1e6:.0f} GB/s written")
print(f"expand + condense + assemble from cell coefficients: {tp:.2f} ms = {n / tp / 1e3:.1f} M cells/s (info bad: {int(info.abs().sum())})")
