"""Generates tests/golden/*.npz: seeded inputs + outputs of the SciPy-LAPACK oracle (dgetrf/dgetrs/dgemm/dgemv, the
entry points the reference calls at src/StaticCondensationMap.jl:179-192 and src/BackwardStaticCondensationMap.jl:91-99)
and of the restated Gridap assembler.  The reference itself (Julia) cannot run in the build container, so these are the
pinned known answers of this repo; regenerate with:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as o  # noqa: E402
from tests.helpers import CONFIGS, DarcyProblem  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    for name in ["C1_hdg_k1_2d", "C2_rth_k1_2d", "C3_hdg_k2_3d", "multifield_2skel"]:
        c = CONFIGS[name]
        plan = o.BlockPlan(c["ndofs"], c["touched"], c["interior"], c["boundary"])
        ncells, cell_start = 6, 424242
        A, b = o.synth_cell_records(plan, cell_start, ncells)
        S, g, info = o.condense_records(plan, A, b)
        x = np.random.default_rng(17).standard_normal((ncells, plan.n_b))
        u, _ = o.backsub_records(plan, A, b, x)
        np.savez_compressed(os.path.join(HERE, f"cells_{name}.npz"), cell_start=cell_start, S=S, g=g, info=info, x=x, u=u,
                            A_checksum=np.array([A.sum(), np.abs(A).sum()]))
    # Darcy HDG 3x2 quads, order 1: ids, CSC pattern, values, rhs, lambda, u of the oracle pipeline
    prob = DarcyProblem((3, 2), 1)
    out = prob.oracle_solve()
    np.savez_compressed(os.path.join(HERE, "darcy_hdg_3x2_k1.npz"), cell_wise_facets=prob.cwf, cell_ids=prob.cell_ids,
                        colptr=out["colptr"], rowval=out["rowval"], nzval=out["nzval"], rhs=out["rhs"], lam=out["lam"],
                        u=out["u"], S=out["S"], g=out["g"])
    # bulk -> skeleton L2 projection dofs (src/GridapAPIExtensions.jl:453-500, A\B per (cell, local facet)): facet mass
    # matrices of P2 on a hex face in the monomial basis (6 x 6, scaled per system) against 30 bulk moments, dgetrf/dgetrs
    rng = np.random.default_rng(23)
    xq, wq = np.polynomial.legendre.leggauss(4)
    xq, wq = 0.5 * (xq + 1), 0.5 * wq
    pts = np.array([(a, b_) for a in xq for b_ in xq]); w = np.array([wa * wb for wa in wq for wb in wq])
    expo = [(0, 0), (0, 1), (1, 0), (0, 2), (1, 1), (2, 0)]
    L = np.stack([pts[:, 0] ** e[0] * pts[:, 1] ** e[1] for e in expo], axis=1)
    M = (L * w[:, None]).T @ L
    areas = rng.uniform(0.01, 2.0, 9)
    Am = M[None] * areas[:, None, None]
    Bm = rng.standard_normal((9, 6, 30))
    Xm, info = o.l2_projection_dofs(Am, Bm)
    np.savez_compressed(os.path.join(HERE, "l2_projection_p2_facets.npz"), A=Am, B=Bm, X=Xm, info=info)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
