"""A mesh that does not fit with resident records: C3 from the coefficient vectors of an affine family on ONE GPU
(records never exist: ghb_condense_assemble_affine_f64 + ghb_backsub_affine_f64).  160^3 = 4 096 000 cells would need
163 GB of records alone."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gridaphybrid_b200 as gh  # noqa: E402

d = int(sys.argv[1]) if len(sys.argv) > 1 else 160
dims = (d, d, d)
ctx = gh.Context(0)
plan = ctx.plan_blocks([30, 4, 36], np.ones((3, 3), bool), [1, 2], [3])
n = int(np.prod(dims))
rng = np.random.default_rng(0)
A0 = torch.empty((1, plan.lenA), dtype=torch.float64, device="cuda"); b0 = torch.empty((1, plan.lenb), dtype=torch.float64, device="cuda")
ctx.synth_fill(plan, 0, 1, A0, b0)
ntab = 7
fam = gh.AffineRecordFamily(np.concatenate([A0.cpu().numpy(), 1e-3 * rng.standard_normal((ntab - 1, plan.lenA))]),
                            np.concatenate([b0.cpu().numpy(), 1e-3 * rng.standard_normal((ntab - 1, plan.lenb))]))
coef = gh.cartesian_coefficients(dims, tuple(1.0 / x for x in dims), "cuda")
t0 = time.perf_counter()
sk = gh.CartesianSkeleton(dims, ctx)
M = gh.FacetFESpace(sk, 6, sk.facet_is_boundary())
asm = gh.SparseMatrixAssembler(M)
colptr, rowval, nnz = asm.symbolic()
torch.cuda.synchronize()
print(f"{n} cells ({d}^3), nnz {nnz}, rows {asm.nrows}; symbolic {time.perf_counter() - t0:.1f} s; "
      f"records would be {n * (plan.lenA + plan.lenb) * 8 / 1e9:.0f} GB; device memory in use {torch.cuda.mem_get_info()[1] / 1e9 - torch.cuda.mem_get_info()[0] / 1e9:.0f} GB")
nz = torch.empty(nnz, dtype=torch.float64, device="cuda"); rhs = torch.empty(asm.nrows, dtype=torch.float64, device="cuda")
info = torch.empty(n, dtype=torch.int32, device="cuda")
ev = lambda: torch.cuda.Event(enable_timing=True)


def timed(f, reps=2):
    f(); torch.cuda.synchronize()
    e0, e1 = ev(), ev(); e0.record()
    for _ in range(reps):
        f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


asm.select()
ta = timed(lambda: fam.condense_assemble(ctx, plan, coef, None, nz, rhs, info))
assert int(info.abs().sum()) == 0
print(f"coefficients -> CSC values + rhs: {ta:.1f} ms = {n / ta / 1e3:.1f} M cells/s   (device memory in use "
      f"{(torch.cuda.mem_get_info()[1] - torch.cuda.mem_get_info()[0]) / 1e9:.0f} GB)")
lam = torch.randn(asm.nrows, dtype=torch.float64, device="cuda")
u = torch.empty((n, plan.n_i), dtype=torch.float64, device="cuda")
tb = timed(lambda: fam.backsub(ctx, plan, coef, lam, None, asm.cell_ids, u, info))
print(f"backward map from coefficients:   {tb:.1f} ms = {n / tb / 1e3:.1f} M cells/s; checksum nz {float(nz.sum()):.6e} u {float(u.sum()):.6e}")
