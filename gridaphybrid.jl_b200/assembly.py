"""Mirror of the assembler call sites of the hot path.

`assemble_matrix_and_vector(SparseMatrixAssembler(M,L), data)` (src/HybridAffineFEOperators.jl:38,46;
src/HybridLinearSolvers.jl:43-44) -> `SparseMatrixCSC{Float64,Int64}` + `Vector{Float64}` with the
indices Julia's `sparse(I,J,V,m,n)` produces.
"""
from __future__ import annotations

import numpy as np
import torch

from .blocks import CondensedCells
from .context import Context


class SparseMatrixCSC:
    """Julia-compatible CSC on the device: 1-based Int64 `colptr` [n+1], `rowval` [nnz], Float64 `nzval`."""

    def __init__(self, m, n, colptr, rowval, nzval):
        self.m, self.n, self.colptr, self.rowval, self.nzval = int(m), int(n), colptr, rowval, nzval

    @property
    def nnz(self):
        return int(self.nzval.numel())

    def to_scipy(self):
        import scipy.sparse as sp
        return sp.csc_matrix((self.nzval.cpu().numpy(), (self.rowval - 1).cpu().numpy(), (self.colptr - 1).cpu().numpy()),
                             shape=(self.m, self.n))


class SparseMatrixAssembler:
    """`SparseMatrixAssembler(M, L)`: owns the cell dof ids of the skeleton space and the cached
    symbolic phase (pattern + gather map) inside the Context."""

    def __init__(self, trial, test=None, ctx: Context | None = None):
        self.trial = trial
        self.test = test if test is not None else trial
        self.ctx = ctx or trial.skeleton.ctx
        self.cell_ids = trial.cell_dof_ids()
        self.nrows = trial.num_free_dofs
        self._pattern = None
        self._pid = -1            # handle of this assembler's symbolic pattern inside the Context

    def select(self):
        """make this assembler's pattern the selected one of the Context (other assemblers may share it)"""
        self.symbolic()
        self.ctx.assemble_select(self._pid)

    def symbolic(self):
        if self._pattern is None:
            ncells, n_b = self.cell_ids.shape
            self.ctx.use_torch_stream()
            nnz = self.ctx.assemble_symbolic(ncells, n_b, self.cell_ids, self.nrows)
            self._pid = self.ctx.assemble_current()
            dev = self.cell_ids.device
            colptr = torch.empty(self.nrows + 1, dtype=torch.int64, device=dev)
            rowval = torch.empty(nnz, dtype=torch.int64, device=dev)
            self.ctx.assemble_pattern(colptr, rowval)
            self._pattern = (colptr, rowval, nnz)
        return self._pattern


def attach_dirichlet(assem: SparseMatrixAssembler):
    """`_attach_dirichlet(matvec, mat, uhd)` (src/HybridAffineFEOperators.jl:41-42): in this layout the
    lift g_K <- g_K - S_K*vals_K is applied inside the numeric assembly; return the Dirichlet values."""
    return assem.trial.dirichlet_values


def assemble_matrix_and_vector(assem: SparseMatrixAssembler, data, dirichlet_values=None):
    """`data` = CondensedCells (cell ids come from the assembler's space, as `_collect_cell_matrix_and_vector`
    pairs them, src/HybridAffineFEOperators.jl:44).  `dirichlet_values` None = no lift (Newton path,
    src/HybridLinearSolvers.jl:37-41)."""
    assert isinstance(data, CondensedCells)
    colptr, rowval, nnz = assem.symbolic()
    dev = assem.cell_ids.device
    nzval = torch.empty(nnz, dtype=torch.float64, device=dev)
    rhs = torch.empty(assem.nrows, dtype=torch.float64, device=dev)
    assem.ctx.use_torch_stream()
    assem.select()
    assert data.S.numel() == assem.cell_ids.shape[0] * assem.cell_ids.shape[1] ** 2, "S does not match the assembler's cells"
    assem.ctx.assemble_numeric(data.S, data.g, dirichlet_values, nzval, rhs)
    return SparseMatrixCSC(assem.nrows, assem.nrows, colptr, rowval, nzval), rhs


class SparseMatrixCSR:
    """`SparseMatrixCSR{1,Float64,Int64}` (SparseMatricesCSR.jl, the reference's Manifest pins it) on the device:
    1-based `rowptr` [m+1], `colval` [nnz] ascending within a row, `nzval`."""

    def __init__(self, m, n, rowptr, colval, nzval):
        self.m, self.n, self.rowptr, self.colval, self.nzval = int(m), int(n), rowptr, colval, nzval

    @property
    def nnz(self):
        return int(self.nzval.numel())

    def to_scipy(self):
        import scipy.sparse as sp
        return sp.csr_matrix((self.nzval.cpu().numpy(), (self.colval - 1).cpu().numpy(), (self.rowptr - 1).cpu().numpy()),
                             shape=(self.m, self.n))


def assemble_matrix_and_vector_csr(assem: SparseMatrixAssembler, data, dirichlet_values=None):
    """CSR hand-off of the same system: the pattern is structurally symmetric, so rowptr/colval are the
    CSC pattern; the values are gathered from the transposed cell blocks.  `data.S` is transposed IN PLACE."""
    assert isinstance(data, CondensedCells)
    colptr, rowval, nnz = assem.symbolic()
    dev = assem.cell_ids.device
    nzval = torch.empty(nnz, dtype=torch.float64, device=dev)
    rhs = torch.empty(assem.nrows, dtype=torch.float64, device=dev)
    assem.ctx.use_torch_stream()
    assem.select()
    assert data.S.numel() == assem.cell_ids.shape[0] * assem.cell_ids.shape[1] ** 2, "S does not match the assembler's cells"
    assem.ctx.assemble_numeric_csr(data.S, data.g, dirichlet_values, nzval, rhs)
    return SparseMatrixCSR(assem.nrows, assem.nrows, colptr, rowval, nzval), rhs


def condense_and_assemble(assem: SparseMatrixAssembler, plan, cells, dirichlet_values=None, nzval=None, rhs=None,
                          info=None):
    """Fused site: `lazy_map(StaticCondensationMap, t)` consumed directly by `assemble_matrix_and_vector`
    (ghb_condense_assemble_f64).  `cells.A/b` may be host arrays: they are streamed."""
    colptr, rowval, nnz = assem.symbolic()
    dev = assem.cell_ids.device
    if nzval is None:
        nzval = torch.empty(nnz, dtype=torch.float64, device=dev)
    if rhs is None:
        rhs = torch.empty(assem.nrows, dtype=torch.float64, device=dev)
    assem.ctx.use_torch_stream()
    assem.select()
    assem.ctx.condense_assemble(plan, len(cells), cells.A, cells.b, dirichlet_values, nzval, rhs, info)
    return SparseMatrixCSC(assem.nrows, assem.nrows, colptr, rowval, nzval), rhs
