"""A/B of the two DMMA condensation kernels (right-looking bottom block in shared memory vs left-looking bottom block in
registers, GHB_DMMA_LL) on the three DMMA shapes: throughput with CUDA events on inputs larger than L2, and the
difference of their results (both are checked against the oracle by tests/test_gpu_parity.py)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gridaphybrid_b200 as gh  # noqa: E402

RTH = np.array([[1, 1, 1], [1, 0, 0], [1, 0, 0]], bool)
SHAPES = {
    "(34,36)": ([30, 4, 36], np.ones((3, 3), bool), 1 << 20),
    "(33,12)": ([24, 9, 12], RTH, 1 << 20),
    "(56,16)": ([40, 16, 16], RTH, 1 << 19),
}
only = os.environ.get("GHB_AB_ONLY")
ctx = gh.Context(0)
ctx.set_option("cw", 0)      # these experiments are about the 4-warps-per-cell kernels
ev = lambda: torch.cuda.Event(enable_timing=True)


def run(plan, n, A, b, env):
    ctx.set_option("dmma_ll", int(env.get("GHB_DMMA_LL", "1")))
    ctx.set_option("ll_ctas", int(env.get("GHB_LL_CTAS", "0")))
    S = torch.empty((n, plan.n_b ** 2), dtype=torch.float64, device="cuda")
    g = torch.empty((n, plan.n_b), dtype=torch.float64, device="cuda")
    info = torch.empty(n, dtype=torch.int32, device="cuda")
    for _ in range(3):
        ctx.condense(plan, n, A, b, S, g, info)
    e0, e1 = ev(), ev()
    e0.record()
    for _ in range(5):
        ctx.condense(plan, n, A, b, S, g, info)
    e1.record()
    torch.cuda.synchronize()
    assert int(info.abs().sum()) == 0
    return n / (e0.elapsed_time(e1) / 5) / 1e3, S, g


def dense_cell(ndofs, touched, Arec, brec):
    """Packed record (touched blocks, block-column-major, each column-major) -> dense n x (n+1) [A | b]."""
    off = np.concatenate(([0], np.cumsum(ndofs)))
    n = off[-1]
    M = np.zeros((n, n + 1), dtype=np.longdouble)
    pos = 0
    for j in range(len(ndofs)):
        for i in range(len(ndofs)):
            if touched[i, j]:
                blk = Arec[pos:pos + ndofs[i] * ndofs[j]].reshape((ndofs[i], ndofs[j]), order="F")
                M[off[i]:off[i + 1], off[j]:off[j + 1]] = blk
                pos += ndofs[i] * ndofs[j]
    M[:, n] = brec
    return M


def schur_longdouble(M, ni):
    """[S | g] = [A22 b2] - A21 A11^-1 [A12 b1] in 80-bit arithmetic (partial pivoting); interior fields come first."""
    M = M.copy()
    for k in range(ni):
        piv = k + int(np.argmax(np.abs(M[k:ni, k])))
        M[[k, piv]] = M[[piv, k]]
        l = M[k + 1:, k] / M[k, k]
        M[k + 1:, k:] -= np.outer(l, M[k, k:])
    return np.asarray(M[ni:, ni:])


print("| shape | variant | M cells/s | max rel diff vs right-looking kernel |")
print("|---|---|---|---|")
for name, (ndofs, touched, n) in SHAPES.items():
    if only and only not in name:
        continue
    plan = ctx.plan_blocks(ndofs, touched, [1, 2], [3])
    A = torch.empty((n, plan.lenA), dtype=torch.float64, device="cuda")
    b = torch.empty((n, plan.lenb), dtype=torch.float64, device="cuda")
    ctx.synth_fill(plan, 0, n, A, b)
    r0, S0, g0 = run(plan, n, A, b, {"GHB_DMMA_LL": "0"})
    print(f"| {name} | right-looking (smem bottom block) | {r0:.2f} | |", flush=True)
    variants = [{"GHB_DMMA_LL": "1"}]
    if name == "(34,36)":
        variants = [{"GHB_DMMA_LL": "1", "GHB_LL_CTAS": str(c)} for c in (8, 7, 6)]
    for env in variants:
        r1, S1, g1 = run(plan, n, A, b, env)
        sc = S0.abs().amax(dim=1, keepdim=True)
        dS = float(((S1 - S0).abs() / sc).max())
        dg = float(((g1 - g0).abs() / g0.abs().amax(dim=1, keepdim=True)).max())
        print(f"| {name} | left-looking {env} | {r1:.2f} | S {dS:.2e}, g {dg:.2e} |", flush=True)
        if dS > 1e-12:
            # which of the two is off?  80-bit reference of the worst cell
            c = int(((S1 - S0).abs() / sc).amax(dim=1).argmax())
            ref = schur_longdouble(dense_cell(ndofs, touched, A[c].cpu().numpy(), b[c].cpu().numpy()), plan.n_i)
            Sr = np.asarray(ref[:, :plan.n_b], dtype=np.float64).flatten(order="F")
            e0 = np.abs(S0[c].cpu().numpy() - Sr).max() / np.abs(Sr).max()
            e1 = np.abs(S1[c].cpu().numpy() - Sr).max() / np.abs(Sr).max()
            print(f"|  | worst cell {c}: error vs 80-bit reference | right-looking {e0:.2e} | left-looking {e1:.2e} |", flush=True)
        del S1, g1
    del A, b, S0, g0
    torch.cuda.empty_cache()
