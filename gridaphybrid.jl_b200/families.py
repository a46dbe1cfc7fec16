"""Element records of an affine family, generated on the device.

On a Cartesian (or any affine) mesh with cell-wise constant coefficients the element matrices that the reference
integrates cell by cell (`lazy_map(::IntegrationMap, ...)`, src/GridapAPIExtensions.jl:442-451) are linear combinations
of a few reference records, `A_K = sum_t coef[K][t] * TA[t]`.  The tables come from the reference's *own* integration on
`ntab` representative cells (`from_representatives`: no finite-element code lives here), the coefficients from the cell
index; `ghb_expand_records_f64` writes the packed records straight into HBM, so they never cross PCIe.
"""
from __future__ import annotations

import numpy as np
import torch

from .blocks import PackedCells
from .context import BlockPlan, Context


class AffineRecordFamily:
    def __init__(self, TA, Tb):
        self.TA = np.ascontiguousarray(TA, dtype=np.float64)      # [ntab][lenA]
        self.Tb = np.ascontiguousarray(Tb, dtype=np.float64)      # [ntab][lenb]
        assert self.TA.shape[0] == self.Tb.shape[0]
        self.ntab = int(self.TA.shape[0])
        self._dev = None

    @classmethod
    def from_representatives(cls, coef_rep, A_rep, b_rep):
        """Tables from `ntab` representative cells: `coef_rep` [ntab][ntab] are their coefficient vectors (must be
        invertible), `A_rep` [ntab][lenA], `b_rep` [ntab][lenb] their packed records as the reference integrates them."""
        C = np.asarray(coef_rep, dtype=np.float64)
        assert C.ndim == 2 and C.shape[0] == C.shape[1], "need as many representative cells as tables"
        return cls(np.linalg.solve(C, np.asarray(A_rep, dtype=np.float64)),
                   np.linalg.solve(C, np.asarray(b_rep, dtype=np.float64)))

    def expand(self, ctx: Context, plan: BlockPlan, coef: torch.Tensor, A=None, b=None):
        """coef [ncells][ntab] (device) -> packed records (device)."""
        ncells = int(coef.shape[0])
        assert coef.shape[1] == self.ntab and self.TA.shape[1] == plan.lenA and self.Tb.shape[1] == plan.lenb
        dev = coef.device
        if self._dev is None or self._dev[0].device != dev:
            self._dev = (torch.as_tensor(self.TA, device=dev), torch.as_tensor(self.Tb, device=dev))
        if A is None:
            A = torch.empty((ncells, plan.lenA), dtype=torch.float64, device=dev)
        if b is None:
            b = torch.empty((ncells, plan.lenb), dtype=torch.float64, device=dev)
        ctx.use_torch_stream()
        ctx.expand_records(plan, ncells, self.ntab, self._dev[0], self._dev[1], coef.contiguous(), A, b)
        return PackedCells(A, b, plan.ndofs, plan.touched)


    def _tables(self, dev):
        if self._dev is None or self._dev[0].device != dev:
            self._dev = (torch.as_tensor(self.TA, device=dev), torch.as_tensor(self.Tb, device=dev))
        return self._dev

    def condense(self, ctx: Context, plan: BlockPlan, coef: torch.Tensor, S=None, g=None, info=None, keep_factors=False):
        """coef [ncells][ntab] (device) -> (S_K, g_K): the records are formed in the loader of the condensation kernel
        and never written to HBM (`ghb_condense_affine_f64`)."""
        ncells = int(coef.shape[0])
        assert coef.shape[1] == self.ntab and self.TA.shape[1] == plan.lenA and self.Tb.shape[1] == plan.lenb
        dev = coef.device
        TA, Tb = self._tables(dev)
        if S is None:
            S = torch.empty((ncells, plan.n_b * plan.n_b), dtype=torch.float64, device=dev)
        if g is None:
            g = torch.empty((ncells, plan.n_b), dtype=torch.float64, device=dev)
        ctx.use_torch_stream()
        ctx.condense_affine(plan, ncells, self.ntab, TA, Tb, coef.contiguous(), S, g, info, keep_factors=keep_factors)
        return S, g

    def backsub(self, ctx: Context, plan: BlockPlan, coef: torch.Tensor, lambda_free, lambda_dirichlet, cell_ids, u=None,
                info=None):
        """backward map u_K = A11^-1 (b1 - A12 lambda_K) from the coefficient vectors (`ghb_backsub_affine_f64`)."""
        ncells = int(coef.shape[0])
        TA, Tb = self._tables(coef.device)
        if u is None:
            u = torch.empty((ncells, plan.n_i), dtype=torch.float64, device=coef.device)
        ctx.use_torch_stream()
        ctx.backsub_affine(plan, ncells, self.ntab, TA, Tb, coef.contiguous(), lambda_free, lambda_dirichlet, cell_ids, u, info)
        return u

    def condense_assemble(self, ctx: Context, plan: BlockPlan, coef: torch.Tensor, dirichlet_vals, nzval, rhs, info=None):
        """coefficients -> CSC values + rhs of the selected symbolic pattern (`ghb_condense_assemble_affine_f64`)."""
        TA, Tb = self._tables(coef.device)
        ctx.use_torch_stream()
        ctx.condense_assemble_affine(plan, int(coef.shape[0]), self.ntab, TA, Tb, coef.contiguous(), dirichlet_vals,
                                     nzval, rhs, info)
        return nzval, rhs


class AffineCells:
    """Lazy cell array of an affine family: what the reference's `lazy_map(::IntegrationMap, ...)` array is to its
    assembler -- cell (A_K, b_K) exists only while it is being condensed.  `lazy_map(StaticCondensationMap(...), cells)`
    condenses it through `ghb_condense_affine_f64` and the operators recover the bulk unknowns through
    `ghb_backsub_affine_f64` (records formed in the loader of the kernels, never in HBM); `materialise` writes the packed
    records out for callers that want them (bit-identical to what the kernels formed)."""

    def __init__(self, family: AffineRecordFamily, coef: torch.Tensor, ndofs, touched, ctx: Context | None = None):
        self.family, self.coef = family, coef.contiguous()
        self.ndofs = [int(x) for x in ndofs]
        nf = len(self.ndofs)
        self.touched = np.asarray(touched, dtype=bool).reshape(nf, nf)
        self.ncells = int(coef.shape[0])
        self.ctx = ctx
        self._packed = None

    def __len__(self):
        return self.ncells

    def materialise(self, ctx: Context, plan: BlockPlan) -> PackedCells:
        if self._packed is None:
            self._packed = self.family.expand(ctx, plan, self.coef)
        return self._packed

    def _need(self):
        assert self._packed is not None, "AffineCells: records not materialised yet (materialise(ctx, plan))"
        return self._packed

    @property
    def A(self):
        return self._need().A

    @property
    def b(self):
        return self._need().b


def cartesian_coefficients(dims, h, device, cell_start=0, ncells=None, extra=None) -> torch.Tensor:
    """Coefficient vectors of the cells of a Cartesian mesh (x fastest), [ncells][1 + 2 D (+ extras)]:
    1, [idx_a == 0] per axis (the low-side facet is a boundary facet: its owner-normal sign flips, e.g.
    test/DarcyHDGTests.jl:125-135), x0_a = idx_a * h_a per axis (loads that are affine in x); `extra` [ncells][m] are
    appended as they are (cell-wise material coefficients)."""
    dims = [int(d) for d in dims]
    D = len(dims)
    n = int(np.prod(dims)) - cell_start if ncells is None else int(ncells)
    c = torch.arange(cell_start, cell_start + n, dtype=torch.int64, device=device)
    cols = [torch.ones(n, dtype=torch.float64, device=device)]
    idx = []
    for a in range(D):
        idx.append(c % dims[a])
        c = c // dims[a]
    cols += [(idx[a] == 0).to(torch.float64) for a in range(D)]
    cols += [idx[a].to(torch.float64) * float(h[a]) for a in range(D)]
    out = torch.stack(cols, dim=1)
    if extra is not None:
        out = torch.cat([out, extra.to(torch.float64).reshape(n, -1)], dim=1)
    return out.contiguous()
