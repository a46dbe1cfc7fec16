"""CPU tests: pin the oracle against the reference's golden vectors / known answers and LAPACK."""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle import oracle as o
from oracle import oracle_c as oc
from tests.helpers import CONFIGS, DarcyProblem, oracle_plan, rel_err_cells


def test_golden_cell_wise_facets_2x1():
    # /root/reference/test/LinearElasticityHDGTests.jl:292  cfids=[[1,2,3,4],[5,6,4,7]]
    assert o.cartesian_cell_wise_facets((2, 1)).tolist() == [[1, 2, 3, 4], [5, 6, 4, 7]]


def test_derived_cell_wise_facets_2x2():
    # SURVEY Appendix A2 derived example
    assert o.cartesian_cell_wise_facets((2, 2)).tolist() == [[1, 2, 3, 4], [5, 6, 4, 7], [2, 8, 9, 10], [6, 11, 10, 12]]


def test_facets_3d_consistency():
    cwf = o.cartesian_cell_wise_facets((3, 2, 2))
    assert cwf.max() == 4 * 2 * 2 + 3 * 3 * 2 + 3 * 2 * 3
    caf = o.cells_around_facets(cwf)
    assert (caf[:, 0] > 0).all() and ((caf[:, 1] == 0) | (caf[:, 1] > caf[:, 0])).all()
    # x-neighbours share x1/x0, y-neighbours y1/y0, z-neighbours z1/z0 (HEX local facets z0,z1,y0,y1,x0,x1)
    assert cwf[0, 5] == cwf[1, 4] and cwf[0, 3] == cwf[3, 2] and cwf[0, 1] == cwf[6, 0]


def test_scalar2arrayblock_golden():
    # /root/reference/test/Scalar2ArrayBlockMapTests.jl:6-19
    rng = np.random.default_rng(1)
    bs = [8, 16]
    A, b = rng.random((24, 24)), rng.random(24)
    Ab, bb = o.scalar2arrayblock(A, b, bs)
    assert np.array_equal(Ab.array[0][0], A[0:8, 0:8]) and np.array_equal(Ab.array[1][0], A[8:24, 0:8])
    assert np.array_equal(Ab.array[0][1], A[0:8, 8:24]) and np.array_equal(Ab.array[1][1], A[8:24, 8:24])
    assert np.array_equal(bb.array[0], b[:8]) and np.array_equal(bb.array[1], b[8:])


def test_sum_facets_golden():
    # /root/reference/test/SumFacetMapTests.jl:10-29: four equal facet blocks sum to 4a
    rng = np.random.default_rng(2)
    a = rng.random((2, 4))
    ab = o.ArrayBlock([a, None, None], [True, False, False])
    abf = o.ArrayBlock([ab, ab, ab, ab], [True] * 4)
    res = o.sum_facets(abf)
    assert np.allclose(res.array[0], 4 * a)
    # :32-109: disjoint facet blocks [f][1,b][1,f] sum + densify == hcat(a,a,a,a) (facet-major order)
    a = rng.random((2, 1))
    facets = []
    for f in range(4):
        inner = o.ArrayBlock([[a if q == f else None for q in range(4)]], np.array([[q == f for q in range(4)]]))
        facets.append(o.ArrayBlock([[inner, None, None]], np.array([[True, False, False]])))
    res = o.densify_innermost(o.sum_facets(o.ArrayBlock(facets, [True] * 4)))
    assert np.allclose(res.array[0][0], np.hstack([a, a, a, a]))


def test_static_condensation_reference_unit_test_shape():
    # /root/reference/test/StaticCondensationMapTests.jl:6-46 (no asserts there; here: known answer)
    rng = np.random.default_rng(3)
    x = rng.random((3, 3))
    y = [[x, x + 3, x + 5], [x + 1, None, None], [x + 2, None, None]]
    touched = np.ones((3, 3), bool)
    touched[1:, 1:] = False
    xv = rng.random(3)
    A = o.ArrayBlock(y, touched)
    b = o.ArrayBlock([xv, xv + 1, xv + 2], [True] * 3)
    S, g, info = o.static_condensation(A, b, [1, 2], [3])
    assert info == 0
    A11 = np.block([[x, x + 3], [x + 1, np.zeros((3, 3))]])
    A12 = np.vstack([x + 5, np.zeros((3, 3))])
    A21 = np.hstack([x + 2, np.zeros((3, 3))])
    S_ref = -A21 @ np.linalg.solve(A11, A12)
    g_ref = (xv + 2) - A21 @ np.linalg.solve(A11, np.concatenate([xv, xv + 1]))
    assert np.allclose(S, S_ref, rtol=1e-10, atol=1e-12) and np.allclose(g, g_ref, rtol=1e-10, atol=1e-12)
    blk, info = o.backward_static_condensation(A, b, g, [1, 2], [3])
    u_ref = np.linalg.solve(A11, np.concatenate([xv, xv + 1]) - A12 @ g)
    assert np.allclose(np.concatenate(blk.array[:2]), u_ref, rtol=1e-9, atol=1e-11)
    assert np.array_equal(blk.array[2], g)


@pytest.mark.parametrize("name", list(CONFIGS))
def test_c_twin_matches_lapack(name):
    """C restatement (own dgetf2/dgetrs/dgemm loops) vs SciPy's LAPACK entry points the reference calls."""
    p = oracle_plan(name)
    A, b = o.synth_cell_records(p, 7, 24)
    S, g, info = o.condense_records(p, A, b)
    S2, g2, info2 = oc.condense(p, A, b)
    assert not info.any() and not info2.any()
    assert rel_err_cells(S2, S) < 1e-13 and rel_err_cells(g2, g) < 1e-13
    x = np.random.default_rng(5).standard_normal((24, p.n_b))
    u, _ = o.backsub_records(p, A, b, x)
    u2, _ = oc.backsub(p, A, b, x)
    assert rel_err_cells(u2, u) < 1e-13
    assert rel_err_cells(oc.condense(p, A, b, nthreads=1)[0], S2) == 0.0  # threading does not change bits


def test_singular_cell_reports_info():
    p = oracle_plan("C1_hdg_k1_2d")
    A, b = o.synth_cell_records(p, 0, 3)
    A[1, :] = 0.0
    _, _, info = o.condense_records(p, A, b)
    _, _, info2 = oc.condense(p, A, b)
    assert info.tolist() == [0, 1, 0] and info2.tolist() == [0, 1, 0]


def test_julia_sparse_semantics():
    """sparse(I,J,V,m,n): sorted columns/rows, duplicates summed, explicit zeros kept."""
    I = np.array([3, 1, 3, 2, 1]); J = np.array([2, 1, 2, 2, 1]); V = np.array([1.0, 2.0, 4.0, 0.0, -2.0])
    colptr, rowval, nzval = o.julia_sparse(I, J, V, 3, 3)
    assert colptr.tolist() == [1, 2, 4, 4] and rowval.tolist() == [1, 2, 3] and nzval.tolist() == [0.0, 0.0, 5.0]
    rng = np.random.default_rng(0)
    I = rng.integers(1, 30, 500); J = rng.integers(1, 30, 500); V = rng.standard_normal(500)
    colptr, rowval, nzval = o.julia_sparse(I, J, V, 29, 29)
    ref = sp.coo_matrix((V, (I - 1, J - 1)), shape=(29, 29)).tocsc()
    ref.sum_duplicates(); ref.sort_indices()
    assert np.array_equal(colptr - 1, ref.indptr) and np.array_equal(rowval - 1, ref.indices)
    assert np.allclose(nzval, ref.data)


def test_facet_dof_numbering_and_glue():
    cwf = o.cartesian_cell_wise_facets((2, 2))
    isb = o.facet_is_boundary(cwf)
    assert isb.tolist() == [True, False, True, False, True, False, True, True, True, False, True, True]
    ids, nfree, ndir = o.facet_dof_ids(isb, 2)
    assert nfree == 8 and ndir == 16
    assert ids[0].tolist() == [-1, -2] and ids[1].tolist() == [1, 2] and ids[3].tolist() == [3, 4]
    cell_ids = o.restrict_facet_dofs_to_skeleton(cwf, ids)
    assert cell_ids[0].tolist() == [-1, -2, 1, 2, -3, -4, 3, 4]
    assert o.generate_cell_is_dirichlet(cell_ids < 0).all()
    glue = o.glue_facet_and_cell_wise_dofs(o.cells_around_facets(cwf), cwf, 2)
    assert glue[1] == (1, 3, 2) and glue[3] == (1, 7, 2) and glue[5] == (2, 3, 2)


@pytest.mark.parametrize("dims,order", [((2, 2), 1), ((4, 3), 1), ((3, 3), 2), ((2, 2, 2), 1), ((3, 2, 2), 2)])
def test_darcy_hdg_exact_solution_oracle(dims, order):
    """The reference's own criterion (test/DarcyHDGTests.jl:142): ||u-uh||_L2 < 1e-12 through
    condense -> lift -> assemble -> solve -> back-substitute on the oracle."""
    prob = DarcyProblem(dims, order)
    out = prob.oracle_solve()
    assert np.allclose(out["lam"].reshape(-1, prob.prob.Nl)[:, 0], -3.14, atol=1e-10)
    assert prob.prob.l2_error_u(out["u"][:, :prob.prob.D * prob.prob.Nu]) < 1e-12


def test_philox_known_answer():
    # Random123 known-answer vectors for philox4x32-10
    r = o.philox4x32(0, 0, 0, 0, 0, 0)
    assert [int(x) for x in r] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    r = o.philox4x32(0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff)
    assert [int(x) for x in r] == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    r = o.philox4x32(0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344, 0xa4093822, 0x299f31d0)
    assert [int(x) for x in r] == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_l2_projection_dofs_oracle():
    """A\\B restated with dgetrf/dgetrs: against numpy.linalg.solve, vector and matrix right-hand sides, singular info."""
    rng = np.random.default_rng(5)
    n, m, nb = 6, 10, 7
    Q = rng.standard_normal((nb, n, n))
    A = Q @ np.transpose(Q, (0, 2, 1)) + 0.5 * np.eye(n)          # SPD like a facet mass matrix
    B = rng.standard_normal((nb, n, m))
    X, info = o.l2_projection_dofs(A, B)
    assert not info.any() and np.allclose(X, np.linalg.solve(A, B), rtol=1e-12, atol=1e-12)
    x, _ = o.l2_projection_dofs(A, B[:, :, 0])
    assert x.shape == (nb, n) and np.allclose(x, X[:, :, 0])
    A[3] = 0.0
    _, info = o.l2_projection_dofs(A, B)
    assert info[3] == 1 and info.sum() == 1


def test_affine_family_tables_and_coefficients_cpu():
    """host side of SURVEY 8f-1 (no GPU): tables from representative cells reproduce every cell record of the Darcy HDG
    family as sum_t coef[K][t] T[t]; slab-wise coefficient vectors equal the slice of the global ones."""
    import torch
    import gridaphybrid_b200 as gh
    from oracle import hdg_darcy
    from tests.helpers import pack_blocks
    dims, order = (4, 4), 1
    D, n = len(dims), dims[0]
    rep = hdg_darcy.DarcyHDG((3,) * D, order, length=3.0 / n)
    Ar, br = pack_blocks(*rep.cell_blocks())
    picks = [[1] * D] + [[0 if a == d else 1 for a in range(D)] for d in range(D)] + [[2 if a == d else 1 for a in range(D)] for d in range(D)]
    cells = [int(sum(idx[a] * 3 ** a for a in range(D))) for idx in picks]
    coefs = np.array([[1.0] + [float(idx[a] == 0) for a in range(D)] + [idx[a] * rep.h[a] for a in range(D)] for idx in picks])
    fam = gh.AffineRecordFamily.from_representatives(coefs, Ar[cells], br[cells])
    full = hdg_darcy.DarcyHDG(dims, order)
    A, b = pack_blocks(*full.cell_blocks())
    coef = gh.cartesian_coefficients(dims, full.h, "cpu").numpy()
    assert coef.shape == (full.ncells, 1 + 2 * D)
    assert np.abs(coef @ fam.TA - A).max() < 1e-13 * np.abs(A).max()
    assert np.abs(coef @ fam.Tb - b).max() < 1e-13 * np.abs(b).max()
    part = gh.cartesian_coefficients(dims, full.h, "cpu", cell_start=5, ncells=7).numpy()
    assert np.array_equal(part, coef[5:12])
    kappa = torch.arange(7, dtype=torch.float64)
    ext = gh.cartesian_coefficients(dims, full.h, "cpu", cell_start=5, ncells=7, extra=kappa).numpy()
    assert ext.shape == (7, 2 + 2 * D) and np.array_equal(ext[:, -1], kappa.numpy())


@pytest.mark.parametrize("shape", [(34, 36), (33, 12), (56, 16)])
def test_left_looking_bottom_block_lane_emulation(shape):
    """tools/emulate_bottom.py replays the register-level index arithmetic of condense_dmma_ll_kernel's bottom block
    (accumulator fragment fed back as the A operand through the column permutation sg, B-fragment rows c0 + sg(2 tig + s),
    partial last panel) lane by lane against numpy, and checks that the B-fragment loads are bank-conflict free."""
    import importlib.util
    import os
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "emulate_bottom.py")
    spec = importlib.util.spec_from_file_location("emulate_bottom", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.run(*shape)            # asserts error < 1e-12 and no NaN; prints the worst bank-conflict degree
    assert mod.SIGMA == [0, 2, 1, 3, 6, 4, 7, 5]


def test_l2_projection_block_overloads_restatement():
    """block overloads of compute_bulk_to_skeleton_l2_projection_dofs (src/GridapAPIExtensions.jl:547-742): result placed
    at [1, b2] of a 1 x nb MatrixBlock; plain array for one right-hand side; VectorBlock of length 1 for several;
    VectorBlocks entry by entry with untouched entries skipped."""
    rng = np.random.default_rng(0)
    nb, n, m = 5, 4, 3
    A = rng.standard_normal((nb, n, n)) + 3 * np.eye(n)
    B = rng.standard_normal((nb, n, m))
    tA = np.zeros((3, 3), bool); tA[0, 1] = True
    tB = np.zeros((3, 3), bool); tB[0, 2] = True
    aa = [[None, A, None], [None] * 3, [None] * 3]; bb = [[None, None, B], [None] * 3, [None] * 3]
    r, t = o.l2_projection_dofs_blocks((aa, tA), (bb, tB))
    assert t.tolist() == [[False, False, True]] and np.allclose(r[0][2], np.linalg.solve(A, B), rtol=1e-12, atol=0)
    t2 = np.zeros((2, 2), bool); t2[0, 0] = True
    x = o.l2_projection_dofs_blocks(([[A, None], [None, None]], t2), ([B[:, :, 0], None], np.array([True, False])))
    assert isinstance(x, np.ndarray) and np.allclose(x, np.linalg.solve(A, B[:, :, :1])[:, :, 0], rtol=1e-12, atol=0)
    r, t = o.l2_projection_dofs_blocks(([[A, None], [None, None]], t2), ([B, None], np.array([True, False])))
    assert t.tolist() == [True] and np.allclose(r[0], np.linalg.solve(A, B), rtol=1e-12, atol=0)
    tf = np.array([True, False])
    r, t = o.l2_projection_dofs_blocks(([(aa, tA), None], tf), ([(bb, tB), None], tf))
    assert t.tolist() == [True, False] and r[1] is None and r[0][1].tolist() == [[False, False, True]]


@pytest.mark.parametrize("seed,nb,shared", [(0, 6, True), (1, 4, False), (2, 9, True)])
def test_c_assembly_equals_julia_sparse_restatement(seed, nb, shared):
    """oracle_c.ora_assemble_coo_csc (Gridap's COO numeric loop + the step-by-step restatement of SparseArrays.sparse!,
    the timed CPU baseline) against the oracle's literal `sparse(I,J,V)` semantics: colptr / rowval bit-equal, values
    bit-equal (duplicates summed in COO order, also with more than two contributions per entry), rhs, Dirichlet ids
    skipped, empty columns."""
    rng = np.random.default_rng(seed)
    nc, nfree = 37, 60 if shared else 37 * nb + 5
    ids = np.stack([rng.choice(np.arange(1, nfree + 1), nb, replace=False) for _ in range(nc)])
    ids[rng.random((nc, nb)) < 0.2] *= -1
    S = rng.standard_normal((nc, nb * nb)); g = rng.standard_normal((nc, nb))
    Sc = [S[c].reshape((nb, nb), order="F") for c in range(nc)]
    c0, r0, z0, b0 = o.assemble_matrix_and_vector(Sc, list(g), ids, nfree)
    c1, r1, z1, b1 = oc.assemble_coo_csc(S, g, ids, nfree)
    assert np.array_equal(c0, c1) and np.array_equal(r0, r1) and np.array_equal(z0, z1) and np.array_equal(b0, b1)
    M = sp.coo_matrix((np.concatenate([[0.0], z1]), (np.concatenate([[0], r1 - 1]), np.concatenate([[0], np.repeat(np.arange(nfree), np.diff(c1))]))), shape=(nfree, nfree))
    assert M.nnz >= len(z1)
