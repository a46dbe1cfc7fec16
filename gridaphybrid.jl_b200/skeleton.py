"""Index semantics of the skeleton for the hot path: cell -> local facet -> global facet -> dof ids.

Only what condensation/assembly/back-substitution need from src/Skeleton.jl / src/SkeletonArrays.jl:
`cell_wise_facets` (src/HybridAffineFEOperators.jl:152-156), the L2 facet dof numbering of the skeleton
space, and RestrictFacetDoFsToSkeleton (src/HybridAffineFEOperators.jl:388-439).
The lazy-array machinery and all geometry (normals, reference maps, VTK) stay with the reference.
"""
from __future__ import annotations

import numpy as np
import torch

from .context import Context


class CartesianSkeleton:
    """Facet topology of a CartesianDiscreteModel `dims` (x fastest), for cells
    [cell_start, cell_start+ncells) -- a slab of the mesh when sharded across GPUs.
    Facet ids are the *global* first-touch ids of the whole mesh (closed form on the device)."""

    def __init__(self, dims, ctx: Context, cell_start: int = 0, ncells: int | None = None):
        self.dims = tuple(int(d) for d in dims)
        self.D = len(self.dims)
        self.ctx = ctx
        tot = int(np.prod(self.dims))
        self.cell_start = int(cell_start)
        self.ncells = tot - self.cell_start if ncells is None else int(ncells)
        self.ncells_global = tot
        self.nlfacets = 2 * self.D
        nf = 0
        for a in range(self.D):
            others = tot // self.dims[a]
            nf += (self.dims[a] + 1) * others
        self.nfacets = nf
        dev = torch.device("cuda", ctx.device)
        self.cell_wise_facets = torch.empty((self.ncells, self.nlfacets), dtype=torch.int64, device=dev)
        ctx.use_torch_stream()
        ctx.cartesian_cell_wise_facets(self.dims, self.cell_start, self.ncells, self.cell_wise_facets)

    def facet_is_boundary(self) -> torch.Tensor:
        """bool [nfacets] (whole mesh): facets with a single cell around them.  Closed form from the
        cell index, so it is also valid for a slab (cells_around_facets, :158-162, never materialised)."""
        dev = self.cell_wise_facets.device
        mask = torch.zeros(self.nfacets, dtype=torch.bool, device=dev)
        c = torch.arange(self.cell_start, self.cell_start + self.ncells, device=dev)
        stride = 1
        for a in range(self.D):
            ia = (c // stride) % self.dims[a]
            lf_low = 2 * (self.D - 1 - a)
            mask[self.cell_wise_facets[ia == 0, lf_low] - 1] = True
            mask[self.cell_wise_facets[ia == self.dims[a] - 1, lf_low + 1] - 1] = True
            stride *= self.dims[a]
        return mask


class FacetFESpace:
    """L2-conforming facet space (the reference's `TestFESpace(Γ, reffe; conformity=:L2, dirichlet_tags=…)`
    / `TrialFESpace(M, g)`), reduced to what the path needs: `ndofs_f` dofs per facet, a Dirichlet mask
    over facets and the Dirichlet values.  Numbering per SURVEY A3."""

    def __init__(self, skeleton: CartesianSkeleton, ndofs_f: int, facet_is_dirichlet: torch.Tensor,
                 dirichlet_values: torch.Tensor | None = None):
        self.skeleton = skeleton
        self.ndofs_f = int(ndofs_f)
        isd = facet_is_dirichlet.to(torch.bool)
        dev = isd.device
        free_rank = torch.cumsum((~isd).to(torch.int64), 0)   # 1-based rank among free facets
        dir_rank = torch.cumsum(isd.to(torch.int64), 0)
        d = torch.arange(1, self.ndofs_f + 1, device=dev, dtype=torch.int64)[None, :]
        free_ids = (free_rank[:, None] - 1) * self.ndofs_f + d
        dir_ids = -((dir_rank[:, None] - 1) * self.ndofs_f + d)
        self.facet_dof_ids = torch.where(isd[:, None], dir_ids, free_ids).contiguous()
        self.num_free_dofs = int(free_rank[-1].item()) * self.ndofs_f
        self.num_dirichlet_dofs = int(dir_rank[-1].item()) * self.ndofs_f
        if dirichlet_values is None:
            dirichlet_values = torch.zeros(max(self.num_dirichlet_dofs, 1), dtype=torch.float64, device=dev)
        self.dirichlet_values = dirichlet_values.to(torch.float64).contiguous()

    def cell_dof_ids(self, free_offset: int = 0, dirichlet_offset: int = 0) -> torch.Tensor:
        """`get_cell_dof_ids` restricted to the cell boundary (RestrictFacetDoFsToSkeleton): int64
        [ncells, nlfacets*ndofs_f]; multi-field offsets as in MultiFieldFESpace."""
        sk = self.skeleton
        ids = self.facet_dof_ids
        if free_offset or dirichlet_offset:
            ids = torch.where(ids > 0, ids + free_offset, ids - dirichlet_offset).contiguous()
        out = torch.empty((sk.ncells, sk.nlfacets * self.ndofs_f), dtype=torch.int64, device=ids.device)
        sk.ctx.use_torch_stream()
        sk.ctx.restrict_facet_dofs(sk.ncells, sk.nlfacets, self.ndofs_f, sk.cell_wise_facets, ids, out)
        return out


class MultiFieldFacetFESpace:
    """`MultiFieldFESpace([trial[i] for i in skeleton_fields])` (src/HybridAffineFEOperators.jl:52-63):
    free dofs of field f are offset by the free dofs of the previous fields; cell ids are concatenated
    field-major -- the order StaticCondensationMap leaves the boundary dofs in."""

    def __init__(self, spaces):
        self.spaces = list(spaces)
        self.skeleton = self.spaces[0].skeleton
        self.num_free_dofs = sum(s.num_free_dofs for s in self.spaces)
        self.num_dirichlet_dofs = sum(s.num_dirichlet_dofs for s in self.spaces)
        self.dirichlet_values = torch.cat([s.dirichlet_values[:s.num_dirichlet_dofs] for s in self.spaces] +
                                          [torch.zeros(1, dtype=torch.float64, device=self.spaces[0].dirichlet_values.device)])
        self.block_sizes = [self.skeleton.nlfacets * s.ndofs_f for s in self.spaces]

    def cell_dof_ids(self) -> torch.Tensor:
        parts, fo, do = [], 0, 0
        for s in self.spaces:
            parts.append(s.cell_dof_ids(fo, do))
            fo += s.num_free_dofs
            do += s.num_dirichlet_dofs
        return torch.cat(parts, dim=1).contiguous()
