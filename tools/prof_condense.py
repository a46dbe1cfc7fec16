"""Small driver for ncu captures: runs the condensation kernel a few times on a synthetic C3 batch."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import gridaphybrid_b200 as gh

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ctx = gh.Context(0)
plan = ctx.plan_blocks([30, 4, 36], np.ones((3, 3), bool), [1, 2], [3])
A = torch.empty((n, plan.lenA), dtype=torch.float64, device="cuda")
b = torch.empty((n, plan.lenb), dtype=torch.float64, device="cuda")
ctx.synth_fill(plan, 0, n, A, b)
S = torch.empty((n, 36 * 36), dtype=torch.float64, device="cuda")
g = torch.empty((n, 36), dtype=torch.float64, device="cuda")
info = torch.empty(n, dtype=torch.int32, device="cuda")
for _ in range(reps):
    ctx.condense(plan, n, A, b, S, g, info)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    ctx.condense(plan, n, A, b, S, g, info)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f"kernel={plan.kernel_name} cells={n} ms={ms:.3f} Mcells/s={n / ms / 1e3:.2f} info_bad={int(info.abs().sum())}")
