"""ncu driver: the fused condensation + scatter kernel on a C3 mesh (64x64x32)."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import gridaphybrid_b200 as gh
from gridaphybrid_b200.distributed import SlabAssembler
ctx = gh.Context(0)
dims = (64, 64, 32)
plan = ctx.plan_blocks([30, 4, 36], np.ones((3, 3), bool), [1, 2], [3])
slab = SlabAssembler(ctx, dims, 6, 0, 1)
n = int(np.prod(dims))
A = torch.empty((n, plan.lenA), dtype=torch.float64, device="cuda"); b = torch.empty((n, plan.lenb), dtype=torch.float64, device="cuda")
ctx.synth_fill(plan, 0, n, A, b)
S = torch.empty((n, 1296), dtype=torch.float64, device="cuda"); g = torch.empty((n, 36), dtype=torch.float64, device="cuda")
info = torch.empty(n, dtype=torch.int32, device="cuda")
nz = torch.empty(slab.nnz, dtype=torch.float64, device="cuda"); rhs = torch.empty(slab.nrows_local, dtype=torch.float64, device="cuda")
for _ in range(4):
    slab.condense_assemble(plan, A, b, S, g, info, nz, rhs)
torch.cuda.synchronize()
print("done", n)
