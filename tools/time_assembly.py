"""Times the numeric assembly kernels (gather_nzval + gather_rhs) alone on the C3 workload (default 128^3 cells)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gridaphybrid_b200 as gh  # noqa: E402

dims = tuple(int(x) for x in sys.argv[1:4]) if len(sys.argv) > 3 else (128, 128, 128)
ctx = gh.Context(0)
sk = gh.CartesianSkeleton(dims, ctx)
M = gh.FacetFESpace(sk, 6, sk.facet_is_boundary())
asm = gh.SparseMatrixAssembler(M)
colptr, rowval, nnz = asm.symbolic()
del rowval
n = asm.cell_ids.shape[0]
S = torch.randn((n, 36 * 36), dtype=torch.float64, device="cuda")
g = torch.randn((n, 36), dtype=torch.float64, device="cuda")
nz = torch.empty(nnz, dtype=torch.float64, device="cuda")
rhs = torch.empty(asm.nrows, dtype=torch.float64, device="cuda")
f = lambda: ctx.assemble_numeric(S, g, M.dirichlet_values, nz, rhs)
for _ in range(3):
    f()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    f()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
gb = (nnz * 10 + n * 1332 * 8 + asm.nrows * 32) / 1e9
print(f"assembly numeric: {ms:.2f} ms for {n} cells, nnz {nnz}: {gb / ms:.2f} TB/s algorithmic")
