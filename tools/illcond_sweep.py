"""Where does the 1e-11 parity bar end?  Cells whose interior block A11 has a prescribed 2-norm condition number
(A11 = Q1 diag(sigma) Q2^T, sigma log-spaced in [1/kappa, 1], random orthogonal Q1, Q2; A12, A21, A22, b ~ N(0,1)): the
condensation kernels of the library against the LAPACK oracle (dgetrf/dgetrs), max over cells of the per-cell relative
Frobenius error of S_K and of g_K.  Two backward-stable LU codes differ by O(kappa * eps) from each other, so the bar is
reachable up to kappa ~ 1e4 whatever the kernel; the table shows how far each kernel is from that line.
usage: python tools/illcond_sweep.py [ncells=256]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gridaphybrid_b200 as gh  # noqa: E402
from oracle import oracle as o  # noqa: E402
from oracle import oracle_c as oc  # noqa: E402
from tests.helpers import CONFIGS, dense_to_record, rel_err_cells  # noqa: E402


def cells_with_condition(op, n, kappa, rng):
    A = np.empty((n, op.lenA)); b = np.empty((n, op.lenb))
    for c in range(n):
        dense = rng.standard_normal((op.n, op.n))
        q1, _ = np.linalg.qr(rng.standard_normal((op.n_i, op.n_i)))
        q2, _ = np.linalg.qr(rng.standard_normal((op.n_i, op.n_i)))
        sig = np.logspace(0, -np.log10(kappa), op.n_i)
        dense[:op.n_i, :op.n_i] = (q1 * sig) @ q2.T
        A[c], b[c] = dense_to_record(op, dense, rng.standard_normal(op.n))
    return A, b


def sweep(ctx, name, n, kappas, force_generic=False):
    cfg = CONFIGS[name]
    op = o.BlockPlan(cfg["ndofs"], cfg["touched"], cfg["interior"], cfg["boundary"])
    ctx.set_option("force_generic", int(force_generic))
    plan = ctx.plan_blocks(cfg["ndofs"], cfg["touched"], cfg["interior"], cfg["boundary"])
    ctx.set_option("force_generic", 0)
    rows = []
    for kappa in kappas:
        rng = np.random.default_rng(int(np.log10(kappa)) + 17)
        A, b = cells_with_condition(op, n, kappa, rng)
        S0, g0, info0 = oc.condense(op, A, b)
        S = np.empty_like(S0); g = np.empty_like(g0); info = np.empty(n, dtype=np.int32)
        ctx.condense(plan, n, A, b, S, g, info)
        assert not info.any() and not info0.any()
        rows.append((kappa, rel_err_cells(S, S0), rel_err_cells(g, g0)))
    return plan.kernel_name, rows


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    ctx = gh.Context(0)
    kappas = [1e1, 1e2, 1e4, 1e6, 1e8, 1e10]
    eps = np.finfo(float).eps
    for name in ("C3_hdg_k2_3d", "C2_rth_k2_2d", "elasticity_k1_2d"):
        for fg in (False, True):
            kn, rows = sweep(ctx, name, n, kappas, fg)
            print(f"{name} [{kn}]")
            for kappa, eS, eg in rows:
                print(f"   cond(A11) = {kappa:7.0e}   err S {eS:9.2e}  err g {eg:9.2e}   (kappa*eps = {kappa * eps:8.1e}, "
                      f"{'inside' if max(eS, eg) < 1e-11 else 'OUTSIDE'} the 1e-11 bar)")
