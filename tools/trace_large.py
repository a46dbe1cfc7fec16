"""Phase timing of condense_large_kernel (build with GHB_NVCC_EXTRA=-DGHB_LTRACE): one cell per SM, CTA 0 prints."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gridaphybrid_b200 as gh
ctx = gh.Context(0)
plan = ctx.plan_blocks([60, 60, 108], np.ones((3, 3), bool), [1, 2], [3])
n = 148
A = torch.empty((n, plan.lenA), dtype=torch.float64, device="cuda"); b = torch.empty((n, plan.lenb), dtype=torch.float64, device="cuda")
ctx.synth_fill(plan, 0, n, A, b)
S = torch.empty((n, plan.n_b ** 2), dtype=torch.float64, device="cuda"); g = torch.empty((n, plan.n_b), dtype=torch.float64, device="cuda")
info = torch.empty(n, dtype=torch.int32, device="cuda")
ctx.condense(plan, n, A, b, S, g, info)
torch.cuda.synchronize()
print("==== second launch (warm) ====", flush=True)
ctx.condense(plan, n, A, b, S, g, info)
torch.cuda.synchronize()
