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
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
