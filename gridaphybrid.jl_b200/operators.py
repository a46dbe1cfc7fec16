"""Mirror of the hybrid operators that call the hot path (API kept, host side thin).

  HybridAffineFEOperator ............ src/HybridAffineFEOperators.jl:1-100
  _compute_hybridizable_from_skeleton_free_dof_values ... :102-150
  HybridFEOperator / HybridBackslashNumericalSetup solve! src/HybridFEOperators.jl:68-141,
                                                          src/HybridLinearSolvers.jl:15-59

What differs from the reference by construction: `weakform` does not build symbolic DomainContributions;
it returns the *integrated* cell-wise block system `PackedCells` (the output of
`_merge_bulk_and_skeleton_contributions` + `_pair_contribution_when_possible`, :25-28), i.e. exactly the
array `StaticCondensationMap` is lazy-mapped over at :338.  The global sparse solve is the caller's
(out of scope, timed separately); SciPy's SuperLU stands in for UMFPACK here.
"""
from __future__ import annotations

import numpy as np
import torch

from .assembly import SparseMatrixAssembler, assemble_matrix_and_vector, attach_dirichlet
from .blocks import PackedCells
from .maps import BackwardStaticCondensationMap, StaticCondensationMap, lazy_map
from .skeleton import MultiFieldFacetFESpace


def _setup_fe_spaces_skeleton_system(trial, skeleton_fields):
    """src/HybridAffineFEOperators.jl:52-63."""
    if len(skeleton_fields) == 1:
        return trial[skeleton_fields[0] - 1]
    return MultiFieldFacetFESpace([trial[i - 1] for i in skeleton_fields])


class AffineFEOperator:
    def __init__(self, trial, test, A, b):
        self.trial, self.test, self.matrix, self.vector = trial, test, A, b


def solve_skeleton(op: AffineFEOperator) -> torch.Tensor:
    """`solve(op.skeleton_op)` (src/HybridAffineFEOperators.jl:74): OUT OF SCOPE global sparse LU."""
    import scipy.sparse.linalg as spla
    x = spla.spsolve(op.matrix.to_scipy().tocsc(), op.vector.cpu().numpy())
    return torch.as_tensor(np.atleast_1d(x), dtype=torch.float64, device=op.vector.device)


class HybridAffineFEOperator:
    """`HybridAffineFEOperator(weakform, trial, test, bulk_fields, skeleton_fields)`.

    trial/test: list indexed by field (1-based ids index it as [f-1]); bulk entries are ints (dofs per
    cell of that L2 bulk field), skeleton entries are FacetFESpace objects.
    """

    def __init__(self, weakform, trial, test, bulk_fields, skeleton_fields):
        self.weakform, self.trial, self.test = weakform, trial, test
        self.bulk_fields, self.skeleton_fields = list(bulk_fields), list(skeleton_fields)
        matvec = weakform()                                                  # :18-28
        assert isinstance(matvec, PackedCells) or hasattr(matvec, "family")       # packed records, or a lazy affine family
        condensed = lazy_map(StaticCondensationMap(self.bulk_fields, self.skeleton_fields), matvec)   # :31
        M = _setup_fe_spaces_skeleton_system(trial, self.skeleton_fields)    # :33
        L = _setup_fe_spaces_skeleton_system(test, self.skeleton_fields)
        # :35-37 _block_skeleton_system_contributions: field-major concatenated ids == the blocked system
        self.assem = SparseMatrixAssembler(M, L)                             # :38
        uhd = attach_dirichlet(self.assem)                                   # :41-42
        A, b = assemble_matrix_and_vector(self.assem, condensed, uhd)        # :44-46
        self.condensed_info = condensed.info
        self.skeleton_op = AffineFEOperator(M, L, A, b)                      # :48

    def solve(self):
        """`solve!(uh, ::LinearFESolver, op)` (:72-100): returns the free dof values of the full space."""
        lh = solve_skeleton(self.skeleton_op)                                # :74
        matvec = self.weakform()                                             # :77-87 (re-evaluated)
        return _compute_hybridizable_from_skeleton_free_dof_values(
            lh, self.skeleton_op.trial.dirichlet_values, self.assem, matvec, self.bulk_fields, self.skeleton_fields)


def _compute_hybridizable_from_skeleton_free_dof_values(lh_free, lh_dirichlet, assem, matvec, bulk_fields,
                                                        skeleton_fields):
    """src/HybridAffineFEOperators.jl:102-150 on the device: gather lambda_K through the cell ids
    (get_cell_dof_values, :113), BackwardStaticCondensationMap (:117-118), scatter to the full space (:134-149)."""
    ctx = assem.ctx
    k = BackwardStaticCondensationMap(bulk_fields, skeleton_fields)
    plan = k.static_condensation.plan(matvec, ctx)
    n = len(matvec)
    dev = assem.cell_ids.device
    u = torch.empty((n, plan.n_i), dtype=torch.float64, device=dev)
    info = torch.empty((n,), dtype=torch.int32, device=dev)
    ctx.use_torch_stream()
    if hasattr(matvec, "family"):                # lazy affine family: the records are formed in the loader here as well
        matvec.family.backsub(ctx, plan, matvec.coef, lh_free, lh_dirichlet, assem.cell_ids, u, info)
    else:
        ctx.backsub(plan, n, matvec.A, matvec.b, lh_free, lh_dirichlet, assem.cell_ids, u, info)
    x = torch.empty(n * plan.n_i + lh_free.numel(), dtype=torch.float64, device=dev)
    ctx.scatter_free_dof_values(plan, n, u, lh_free, x)
    return x


class HybridFEOperator:
    """Nonlinear operator shell (src/HybridFEOperators.jl:68-141): `jacobian_and_residual(x)` must return
    the cell-wise block system PackedCells (A_K = dR_K/dx_K, b_K = -R_K) for the current iterate; the
    linear solve of each Newton step is `hybrid_backslash_solve` below."""

    def __init__(self, jacobian_and_residual, trial, test, bulk_fields, skeleton_fields):
        self.jacobian_and_residual, self.trial, self.test = jacobian_and_residual, trial, test
        self.bulk_fields, self.skeleton_fields = list(bulk_fields), list(skeleton_fields)
        M = _setup_fe_spaces_skeleton_system(trial, self.skeleton_fields)
        self.assem = SparseMatrixAssembler(M, M)     # pattern cached across Newton iterations


def hybrid_backslash_solve(op: HybridFEOperator, matvec: PackedCells) -> torch.Tensor:
    """`solve!(x, ns::HybridBackslashNumericalSetup, b)` (src/HybridLinearSolvers.jl:15-59): condense ->
    assemble (NO Dirichlet lift, :37-41) -> A\\b (:45) -> back-substitute with zero Dirichlet correction (:47-57)."""
    condensed = lazy_map(StaticCondensationMap(op.bulk_fields, op.skeleton_fields), matvec)
    A, b = assemble_matrix_and_vector(op.assem, condensed, None)
    x_skel = solve_skeleton(AffineFEOperator(op.assem.trial, op.assem.test, A, b))
    zeros = torch.zeros_like(op.assem.trial.dirichlet_values)
    return _compute_hybridizable_from_skeleton_free_dof_values(x_skel, zeros, op.assem, matvec, op.bulk_fields,
                                                               op.skeleton_fields)
