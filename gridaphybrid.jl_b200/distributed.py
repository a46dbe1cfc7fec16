"""Multi-GPU sharding of the hot path: one process per GPU, cells partitioned in contiguous slabs along
the slowest Cartesian axis.

Condensation and back-substitution are embarrassingly parallel per cell.  Assembly needs ONE exchange:
a facet is owned by the slab of its first (lowest-id) cell, so every cut plane belongs to the lower
slab; the upper slab sends, per cut-plane facet, the `ndofs_f` columns of S_K (and lifted entries of
g_K) of its bottom-layer cells -- `ghb_pack_cut_plane_f64` + one NCCL send/recv pair per cut.  Because
global facet ids are first-touch ordered and a slab's cells are contiguous, each slab owns a contiguous
range of facet ids, hence of free dofs, hence of CSC *columns*: the per-rank `colptr` segments
concatenate directly into the reference's global `SparseMatrixCSC`.
The second exchange (lambda for the backward step) is a halo receive of the cut-plane dofs, or an
all-gather of the owned lambda ranges.

Index math below is closed form for the boundary condition every reference test uses (all boundary
facets Dirichlet, test/DarcyHDGTests.jl:45) and works on CPU or CUDA tensors (it is shared with the
gloo tests).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from ._lib import GHB_EUNSUPPORTED, GhbError


class SlabLayout:
    """Closed-form global numbering for a z-slab partition of a Cartesian mesh `gdims` (x fastest)."""

    def __init__(self, gdims, ndofs_f, rank, world):
        self.gdims = tuple(int(d) for d in gdims)
        self.D = len(self.gdims)
        assert self.gdims[-1] % world == 0, "the slowest axis must be divisible by the number of slabs"
        self.ndofs_f, self.rank, self.world = int(ndofs_f), int(rank), int(world)
        self.stride = [1]
        for d in self.gdims:
            self.stride.append(self.stride[-1] * d)
        self.ncells_global = self.stride[-1]
        self.layer = self.stride[-2]
        self.ncells = self.ncells_global // world
        self.cell_start = rank * self.ncells
        self.nghost = self.layer if rank < world - 1 else 0
        self.nlfacets = 2 * self.D
        self.n_b = self.nlfacets * self.ndofs_f
        nf = self._free_before_scalar
        self.nrows_global = nf(self.ncells_global) * self.ndofs_f
        self.col_begin = nf(self.cell_start) * self.ndofs_f + 1
        self.col_end = nf(self.cell_start + self.ncells) * self.ndofs_f + 1
        self.nrows_local = self.col_end - self.col_begin

    # number of cells c' < c whose index along axis a is 0 / dims[a]-1 (tensor or int)
    def _count_low(self, c, a):
        hi, rem = c // self.stride[a + 1], c % self.stride[a + 1]
        return hi * self.stride[a] + (torch.clamp(rem, max=self.stride[a]) if torch.is_tensor(rem) else min(rem, self.stride[a]))

    def _count_high(self, c, a):
        hi, rem = c // self.stride[a + 1], c % self.stride[a + 1]
        t = rem - (self.gdims[a] - 1) * self.stride[a]
        t = torch.clamp(t, min=0, max=self.stride[a]) if torch.is_tensor(t) else max(0, min(t, self.stride[a]))
        return hi * self.stride[a] + t

    def _free_before_scalar(self, c):
        """free (interior) facets first touched by cells < c:  D*c - sum_a count_high(c, a)."""
        return self.D * c - sum(self._count_high(c, a) for a in range(self.D))

    def cell_dof_ids(self, cells: torch.Tensor) -> torch.Tensor:
        """Global cell boundary ids [len(cells), n_b] (1-based; Dirichlet negative) of arbitrary cells:
        the composition of first-touch facet ids (A2), L2 facet dof numbering (A3) and
        RestrictFacetDoFsToSkeleton (src/HybridAffineFEOperators.jl:405-434)."""
        D, nf = self.D, self.ndofs_f
        c = cells.to(torch.int64)
        d = torch.arange(1, nf + 1, dtype=torch.int64, device=c.device)[None, :]

        def ranks(cc):
            """per cell: free rank base, dirichlet rank base, idx per axis"""
            free0 = D * cc
            dir0 = torch.zeros_like(cc)
            idx = []
            for a in range(D):
                ch, cl = self._count_high(cc, a), self._count_low(cc, a)
                free0 = free0 - ch
                dir0 = dir0 + ch + cl
                idx.append((cc // self.stride[a]) % self.gdims[a])
            return free0, dir0, idx

        def high_facet_id(cc, axis):
            """signed facet rank (1-based; negative = Dirichlet rank) of the high-side facet of cc along axis"""
            free0, dir0, idx = ranks(cc)
            frank, drank = free0.clone(), dir0.clone()
            out = torch.zeros_like(cc)
            for a in range(D - 1, -1, -1):          # local facet order: axis D-1 low, high, ..., axis 0 low, high
                drank = drank + (idx[a] == 0)        # low-side new facet is always a boundary facet
                is_b = idx[a] == self.gdims[a] - 1
                frank = frank + (~is_b)
                drank = drank + is_b
                if a == axis:
                    out = torch.where(is_b, -drank, frank)
            return out

        def low_boundary_id(cc, axis):
            _, dir0, idx = ranks(cc)
            drank = dir0.clone()
            out = torch.zeros_like(cc)
            for a in range(D - 1, -1, -1):
                drank = drank + (idx[a] == 0)
                if a == axis:
                    out = -drank
                drank = drank + (idx[a] == self.gdims[a] - 1)
            return out

        cols = []
        for lf in range(2 * D):
            axis, side = D - 1 - lf // 2, lf % 2
            if side == 1:
                fr = high_facet_id(c, axis)
            else:
                ia = (c // self.stride[axis]) % self.gdims[axis]
                nb = high_facet_id(torch.clamp(c - self.stride[axis], min=0), axis)
                fr = torch.where(ia == 0, low_boundary_id(c, axis), nb)
            ids = torch.where((fr > 0)[:, None], (fr - 1)[:, None] * nf + d, -((-fr - 1)[:, None] * nf + d))
            cols.append(ids)
        return torch.cat(cols, dim=1).contiguous()

    def slab_cell_ids(self, device) -> torch.Tensor:
        """ids of the slab's own cells followed by its ghost cells (bottom layer of the slab above)."""
        c = torch.arange(self.cell_start, self.cell_start + self.ncells + self.nghost, dtype=torch.int64, device=device)
        return self.cell_dof_ids(c)


def exchange_cut_plane(send_down: torch.Tensor | None, recv_from_up: torch.Tensor | None, rank: int, world: int,
                       group=None):
    """Collective 1: every slab r>0 sends its packed bottom layer to r-1; every slab r<P-1 receives the
    packed bottom layer of r+1.  One batched isend/irecv (NCCL send/recv inside one group)."""
    ops = []
    if rank > 0 and send_down is not None:
        ops.append(dist.P2POp(dist.isend, send_down, rank - 1, group))
    if rank < world - 1 and recv_from_up is not None:
        ops.append(dist.P2POp(dist.irecv, recv_from_up, rank + 1, group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()


def halo_lambda(lam_owned: torch.Tensor, layout: SlabLayout, group=None) -> torch.Tensor:
    """Collective 2 (all-gather variant): every slab contributes its owned lambda range; returns the
    global free-dof vector (what `get_cell_dof_values(lh, dK)` indexes, src/HybridAffineFEOperators.jl:113)."""
    if layout.world == 1:
        return lam_owned
    sizes = [None] * layout.world
    dist.all_gather_object(sizes, int(lam_owned.numel()), group=group)
    m = max(sizes)                      # owned ranges differ by the boundary planes: pad to a common size
    padded = torch.zeros(m, dtype=lam_owned.dtype, device=lam_owned.device)
    padded[:lam_owned.numel()] = lam_owned
    out = [torch.empty(m, dtype=lam_owned.dtype, device=lam_owned.device) for _ in sizes]
    dist.all_gather(out, padded, group=group)
    return torch.cat([o[:s] for o, s in zip(out, sizes)])


class SlabAssembler:
    """Per-rank assembler of the owned CSC columns (device path; needs the CUDA library)."""

    def __init__(self, ctx, gdims, ndofs_f, rank, world, dirichlet_values=None, group=None):
        self.ctx, self.group = ctx, group
        self.layout = L = SlabLayout(gdims, ndofs_f, rank, world)
        dev = torch.device("cuda", ctx.device)
        self.cell_ids = L.slab_cell_ids(dev)
        self.dirichlet_values = dirichlet_values
        ctx.use_torch_stream()
        self.nnz = ctx.assemble_symbolic_slab(L.ncells, L.nghost, L.ndofs_f, L.n_b, self.cell_ids, L.nrows_global,
                                              L.col_begin, L.col_end)
        self._pid = ctx.assemble_current()
        self.nrows_local = L.nrows_local
        stride = L.n_b * L.ndofs_f + L.ndofs_f
        self.send_buf = torch.empty((L.layer, stride), dtype=torch.float64, device=dev) if rank > 0 else None
        self.ghost = torch.empty((L.nghost, stride), dtype=torch.float64, device=dev) if L.nghost else None

    def pattern(self):
        dev = self.cell_ids.device
        colptr = torch.empty(self.nrows_local + 1, dtype=torch.int64, device=dev)
        rowval = torch.empty(self.nnz, dtype=torch.int64, device=dev)
        self.ctx.assemble_select(self._pid)
        self.ctx.assemble_pattern(colptr, rowval)
        return colptr, rowval

    def pack(self, S, g):
        L = self.layout
        if self.send_buf is not None:
            self.ctx.pack_cut_plane(L.layer, L.n_b, L.ndofs_f, S, g, self.cell_ids, self.dirichlet_values, self.send_buf)

    def assemble(self, S, g, nzval, rhs, exchange=None):
        """numeric phase for the owned columns: pack -> cut-plane exchange -> owner-computes gather."""
        L = self.layout
        self.ctx.use_torch_stream()
        self.ctx.assemble_select(self._pid)
        if L.world > 1:
            self.pack(S, g)
            if exchange is not None:
                exchange(self.send_buf, self.ghost, L.rank, L.world, self.group)
            elif getattr(self.ctx, "comm_size", 1) == L.world:
                self.ctx.exchange_cut_plane(self.send_buf, self.ghost)      # NCCL behind the C ABI (ghb_comm_init)
            else:
                exchange_cut_plane(self.send_buf, self.ghost, L.rank, L.world, self.group)
        self.ctx.assemble_numeric_slab(S, g, self.ghost, self.dirichlet_values, nzval, rhs)

    def _exchange(self, exchange=None):
        L = self.layout
        if exchange is not None:
            exchange(self.send_buf, self.ghost, L.rank, L.world, self.group)
        elif getattr(self.ctx, "comm_size", 1) == L.world:
            self.ctx.exchange_cut_plane(self.send_buf, self.ghost)          # NCCL behind the C ABI (ghb_comm_init)
        else:
            exchange_cut_plane(self.send_buf, self.ghost, L.rank, L.world, self.group)

    def condense_assemble(self, plan, A, b, S, g, info, nzval, rhs, exchange=None, events=None):
        """fused step of a slab: the condensation kernel scatters S_K into nzval itself (no second pass over S); S_K is
        kept only for Dirichlet cells and for the bottom layer that feeds the cut-plane exchange"""
        L = self.layout
        self.ctx.use_torch_stream()
        self.ctx.assemble_select(self._pid)
        keep = L.layer if (L.world > 1 and L.rank > 0) else 0
        nzval.zero_()
        if events:
            events[0].record()                               # brackets the fused kernel alone (bench roofline)
        self.ctx.condense_scatter_slab(plan, L.ncells, A, b, S, g, info, nzval, keep, zero_nzval=False)
        if events:
            events[1].record()
        if L.world > 1:
            self.pack(S, g)
            self._exchange(exchange)
        self.ctx.assemble_finish_slab(S, g, self.ghost, self.dirichlet_values, nzval, rhs)

    def condense_assemble_affine(self, plan, family, coef, S, g, info, nzval, rhs, exchange=None):
        """fused step of a slab from the coefficient vectors of an affine family (`ghb_condense_scatter_slab_affine_f64`): the
        records of the slab's cells are formed in the loader of the condensation kernel and never exist in HBM"""
        L = self.layout
        self.ctx.use_torch_stream()
        self.ctx.assemble_select(self._pid)
        keep = L.layer if (L.world > 1 and L.rank > 0) else 0
        TA, Tb = family._tables(coef.device)
        try:
            self.ctx.condense_scatter_slab_affine(plan, L.ncells, family.ntab, TA, Tb, coef.contiguous(), S, g, info, nzval, keep)
        except GhbError as e:
            if e.code != GHB_EUNSUPPORTED:
                raise
            # plans without a cell-warp kernel that can stage the tables: records written out first (same results)
            cells = family.expand(self.ctx, plan, coef)
            return self.condense_assemble(plan, cells.A, cells.b, S, g, info, nzval, rhs, exchange=exchange)
        if L.world > 1:
            self.pack(S, g)
            self._exchange(exchange)
        self.ctx.assemble_finish_slab(S, g, self.ghost, self.dirichlet_values, nzval, rhs)

    def allgather_lambda(self, lam_owned):
        """collective 2 through the C ABI (grouped ncclBroadcast, no object gather): the global free-dof vector"""
        L = self.layout
        if L.world == 1:
            return lam_owned
        counts = [SlabLayout(L.gdims, L.ndofs_f, r, L.world).nrows_local for r in range(L.world)]
        out = torch.empty(sum(counts), dtype=torch.float64, device=lam_owned.device)
        self.ctx.use_torch_stream()
        self.ctx.allgather_lambda(lam_owned, counts, out)
        return out
