"""Shared test helpers: oracle-side pipelines (CPU) used as the checker for the CUDA path."""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from oracle import hdg_darcy
from oracle import oracle as o

_HENCKY_TOUCHED = np.ones((8, 8), bool)
for _i, _j in [(0, 1), (1, 0), (2, 4), (4, 2), (6, 7), (7, 6), (0, 7), (7, 0)]:
    _HENCKY_TOUCHED[_i, _j] = False

# named configurations of BASELINE.json / SURVEY section 8 (ndofs per field, touched, interior, boundary)
CONFIGS = {
    "C1_hdg_k1_2d": dict(ndofs=[6, 1, 8], touched=np.ones((3, 3), bool), interior=[1, 2], boundary=[3]),
    "C2_rth_k1_2d": dict(ndofs=[12, 4, 8], touched=np.array([[1, 1, 1], [1, 0, 0], [1, 0, 0]], bool),
                         interior=[1, 2], boundary=[3]),
    "C2_rth_k2_2d": dict(ndofs=[24, 9, 12], touched=np.array([[1, 1, 1], [1, 0, 0], [1, 0, 0]], bool),
                         interior=[1, 2], boundary=[3]),
    "C2_rth_k3_2d": dict(ndofs=[40, 16, 16], touched=np.array([[1, 1, 1], [1, 0, 0], [1, 0, 0]], bool),
                         interior=[1, 2], boundary=[3]),
    "C3_hdg_k2_3d": dict(ndofs=[30, 4, 36], touched=np.ones((3, 3), bool), interior=[1, 2], boundary=[3]),
    # shapes the reference's own tests / SURVEY section 9 name besides the BASELINE configs: equal-order Darcy HDG k=2 on
    # hexes (40,36) and elasticity HDG k=1 on quads (21,16) (test/LinearElasticityHDGTests.jl:86-94)
    "hdg_equal_order_3d": dict(ndofs=[30, 10, 36], touched=np.ones((3, 3), bool), interior=[1, 2], boundary=[3]),
    "elasticity_k1_2d": dict(ndofs=[9, 12, 16], touched=np.ones((3, 3), bool), interior=[1, 2], boundary=[3]),
    "multifield_2skel": dict(ndofs=[4, 4, 1, 1, 4, 4],
                             touched=np.array([[1, 0, 1, 0, 1, 0], [0, 1, 0, 1, 0, 1], [1, 0, 0, 0, 0, 0],
                                               [0, 1, 0, 0, 0, 0], [1, 0, 0, 0, 0, 0], [0, 1, 0, 0, 0, 0]], bool),
                             interior=[1, 2, 3, 4], boundary=[5, 6]),
    # 2-D shapes of the reference's own tests that only the shape-generic kernel covers: Hencky HDG k=1 on quads (45,24)
    # with two skeleton fields (test/StressAssistedHenckyHDGTests.jl:134-157) and RT-H k=0 (5,4) (test/DarcyRTHTests.jl:80)
    "hencky_k1_2d": dict(ndofs=[6, 6, 3, 9, 9, 12, 8, 16], touched=_HENCKY_TOUCHED, interior=[1, 2, 3, 4, 5, 6],
                         boundary=[7, 8]),
    "rth_k0_2d": dict(ndofs=[4, 1, 4], touched=np.array([[1, 1, 1], [1, 0, 0], [1, 0, 0]], bool), interior=[1, 2],
                      boundary=[3]),
    "odd_shapes": dict(ndofs=[5, 3, 7], touched=np.ones((3, 3), bool), interior=[3, 1], boundary=[2]),
    # BASELINE.json configs 4 and 5: elasticity HDG k=2 on 3-D hexes (sigma 6*10, u 3*20 ||
    # u-hat 6*(3*6)) and Hencky HDG k=1 (six bulk fields || two skeleton fields); the Hencky mask leaves the
    # skeleton-skeleton cross blocks and a few bulk pairs untouched to exercise the zero blocks at this size
    "C4_elasticity_k2_3d": dict(ndofs=[60, 60, 108], touched=np.ones((3, 3), bool), interior=[1, 2], boundary=[3]),
    "C5_hencky_k1_3d": dict(ndofs=[12, 12, 4, 24, 24, 30, 18, 54], touched=_HENCKY_TOUCHED,
                            interior=[1, 2, 3, 4, 5, 6], boundary=[7, 8]),
}


def condensed_dofs(op):
    """(field index 0-based, local dof) of every row of the condensed order: interior fields first, then boundary."""
    return [(f - 1, l) for f in list(op.interior) + list(op.boundary) for l in range(op.ndofs[f - 1])]


def dense_to_record(op, dense, bvec):
    """dense [n, n] matrix and [n] vector in CONDENSED order -> packed record (touched blocks only)."""
    dofs = condensed_dofs(op)
    A = np.zeros(op.lenA)
    b = np.zeros(op.lenb)
    for r, (fr, lr) in enumerate(dofs):
        b[op.field_offset[fr] + lr] = bvec[r]
        for c, (fc, lc) in enumerate(dofs):
            o_ = op.block_offset[fr, fc]
            if o_ >= 0:
                A[o_ + lr + lc * op.ndofs[fr]] = dense[r, c]
    return A, b


def zero_interior_column(op, Arec, col):
    """zero interior column `col` (condensed numbering) of a packed record, rows of the interior fields only."""
    dofs = condensed_dofs(op)
    fc, lc = dofs[col]
    for f in op.interior:
        o_ = op.block_offset[f - 1, fc]
        if o_ >= 0:
            Arec[o_ + lc * op.ndofs[f - 1]: o_ + (lc + 1) * op.ndofs[f - 1]] = 0.0


def oracle_plan(name):
    c = CONFIGS[name]
    return o.BlockPlan(c["ndofs"], c["touched"], c["interior"], c["boundary"])


def rel_err_cells(x, ref):
    """max over cells of ||x_K - ref_K||_F / ||ref_K||_F (the parity norm of BASELINE.md)."""
    x = np.asarray(x).reshape(len(ref), -1)
    ref = np.asarray(ref).reshape(len(ref), -1)
    num = np.linalg.norm(x - ref, axis=1)
    den = np.maximum(np.linalg.norm(ref, axis=1), 1e-300)
    return float(np.max(num / den))


def pack_blocks(mats, vecs, touched):
    """numpy twin of PackedCells.from_blocks (block-column-major, each block col-major)."""
    nf = len(vecs)
    nc = vecs[0].shape[0]
    parts = []
    for j in range(nf):
        for i in range(nf):
            if touched[i, j]:
                parts.append(np.transpose(mats[i][j], (0, 2, 1)).reshape(nc, -1))
    return np.ascontiguousarray(np.concatenate(parts, axis=1)), np.ascontiguousarray(np.concatenate(vecs, axis=1))


class DarcyProblem:
    """Darcy HDG (test/DarcyHDGTests.jl shape) on a Cartesian mesh with all boundary facets Dirichlet."""

    def __init__(self, dims, order=1):
        self.prob = hdg_darcy.DarcyHDG(dims, order)
        self.dims = tuple(dims)
        self.cwf = o.cartesian_cell_wise_facets(self.dims)
        self.is_dir = o.facet_is_boundary(self.cwf)
        self.fids, self.nfree, self.ndir = o.facet_dof_ids(self.is_dir, self.prob.Nl)
        self.cell_ids = o.restrict_facet_dofs_to_skeleton(self.cwf, self.fids)
        self.dir_vals = self.prob.dirichlet_values(int(self.is_dir.sum()))
        mats, vecs, touched = self.prob.cell_blocks()
        self.mats, self.vecs, self.touched = mats, vecs, touched
        self.A, self.b = pack_blocks(mats, vecs, touched)
        self.plan = o.BlockPlan(self.prob.ndofs, touched, [1, 2], [3])

    def oracle_solve(self):
        """condense -> lift -> assemble -> sparse solve -> back-substitute, all on the oracle."""
        p = self.plan
        S, g, info = o.condense_records(p, self.A, self.b)
        assert not info.any()
        nc = len(S)
        Sc = [S[c].reshape((p.n_b, p.n_b), order="F") for c in range(nc)]
        gl = [o.attach_dirichlet(Sc[c], g[c], self.cell_ids[c], self.dir_vals) for c in range(nc)]
        colptr, rowval, nzval, rhs = o.assemble_matrix_and_vector(Sc, gl, self.cell_ids, self.nfree)
        Amat = sp.csc_matrix((nzval, rowval - 1, colptr - 1), shape=(self.nfree, self.nfree))
        lam = np.atleast_1d(spla.spsolve(Amat, rhs))
        xk = o.cell_dof_values(lam, self.dir_vals, self.cell_ids)
        u, info = o.backsub_records(p, self.A, self.b, xk)
        return dict(S=S, g=g, colptr=colptr, rowval=rowval, nzval=nzval, rhs=rhs, lam=lam, u=u)
