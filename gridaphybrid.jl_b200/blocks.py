"""Batched block containers: the array-level view of Gridap's cell-wise `(MatrixBlock, VectorBlock)`.

The reference hands `StaticCondensationMap` one cell at a time (`evaluate!`), each cell being an
`ArrayBlock` (`array` + `touched`) of dense matrices (src/StaticCondensationMap.jl:41-70).  The B200
path intercepts the *array-level* `lazy_map` site instead, so the natural object is the whole cell
array: `PackedCells` = packed records in HBM (layout in include/ghb.h / DESIGN.md).
"""
from __future__ import annotations

import numpy as np
import torch


class ArrayBlock:
    """Mirror of Gridap.Fields.ArrayBlock for ONE cell: `array` (nested lists) + `touched` mask."""

    def __init__(self, array, touched):
        self.array = array
        self.touched = np.asarray(touched, dtype=bool)

    @property
    def shape(self):
        return self.touched.shape


MatrixBlock = ArrayBlock
VectorBlock = ArrayBlock


def block_sizes(A: ArrayBlock):
    """`_compute_brs_bcs` (src/StaticCondensationMap.jl:72-84)."""
    nr, nc = A.touched.shape
    brs, bcs = [None] * nr, [None] * nc
    for j in range(nc):
        for i in range(nr):
            if A.touched[i, j]:
                brs[i], bcs[j] = A.array[i][j].shape
    return brs, bcs


class PackedCells:
    """Cell array of `(A_K, b_K)` in the packed batch format.

    A: float64 [ncells, lenA], b: float64 [ncells, lenb] (torch CUDA tensors, or CPU tensors / numpy
    arrays which the C library stages itself), `ndofs` per field, `touched` [nfields, nfields].
    """

    def __init__(self, A, b, ndofs, touched):
        self.A, self.b = A, b
        self.ndofs = [int(x) for x in ndofs]
        nf = len(self.ndofs)
        self.touched = np.asarray(touched, dtype=bool).reshape(nf, nf)
        self.ncells = int(A.shape[0])

    def __len__(self):
        return self.ncells

    @staticmethod
    def layout(ndofs, touched):
        """Offsets of the touched blocks inside a record: block-column-major, each block col-major."""
        nf = len(ndofs)
        touched = np.asarray(touched, dtype=bool).reshape(nf, nf)
        off = -np.ones((nf, nf), dtype=np.int64)
        o = 0
        for j in range(nf):
            for i in range(nf):
                if touched[i, j]:
                    off[i, j] = o
                    o += ndofs[i] * ndofs[j]
        return off, o

    @classmethod
    def from_blocks(cls, mat_blocks, vec_blocks, touched, device=None):
        """Pack batched blocks: mat_blocks[i][j] float64 [ncells, r_i, c_j] (row, col) or None;
        vec_blocks[i] float64 [ncells, r_i]."""
        nf = len(vec_blocks)
        touched = np.asarray(touched, dtype=bool).reshape(nf, nf)
        ndofs = [int(v.shape[1]) for v in vec_blocks]
        off, lenA = cls.layout(ndofs, touched)
        ncells = int(vec_blocks[0].shape[0])
        dev = device if device is not None else torch.as_tensor(vec_blocks[0]).device
        A = torch.empty((ncells, lenA), dtype=torch.float64, device=dev)
        for j in range(nf):
            for i in range(nf):
                if touched[i, j]:
                    blk = torch.as_tensor(mat_blocks[i][j], dtype=torch.float64).to(dev)
                    assert blk.shape == (ncells, ndofs[i], ndofs[j])
                    o = int(off[i, j])
                    # column-major inside the record: element (r,c) at o + r + c*r_i
                    A[:, o:o + ndofs[i] * ndofs[j]] = blk.transpose(1, 2).reshape(ncells, -1)
        b = torch.cat([torch.as_tensor(v, dtype=torch.float64).to(dev) for v in vec_blocks], dim=1).contiguous()
        return cls(A, b, ndofs, touched)

    @classmethod
    def from_cells(cls, cells, device=None):
        """Pack a Python list of per-cell `(MatrixBlock, VectorBlock)` tuples (the reference's element type)."""
        A0, b0 = cells[0]
        nf = len(b0.array)
        ndofs, _ = block_sizes(A0)
        mats = [[None] * nf for _ in range(nf)]
        for i in range(nf):
            for j in range(nf):
                if A0.touched[i, j]:
                    mats[i][j] = torch.as_tensor(np.stack([np.asarray(c[0].array[i][j], dtype=np.float64) for c in cells]))
        vecs = [torch.as_tensor(np.stack([np.asarray(c[1].array[i], dtype=np.float64) for c in cells])) for i in range(nf)]
        return cls.from_blocks(mats, vecs, A0.touched, device=device)


class CondensedCells:
    """Result of `lazy_map(StaticCondensationMap, t)`: S [ncells, n_b*n_b] (column-major per cell) and
    g [ncells, n_b].  Indexing returns `(S_K, g_K)` views that alias the batch buffer -- the same
    'valid until the next evaluation' contract as the reference's cache aliasing
    (src/StaticCondensationMap.jl:195)."""

    def __init__(self, S, g, info, n_b, plan):
        self.S, self.g, self.info, self.n_b, self.plan = S, g, info, n_b, plan

    def __len__(self):
        return int(self.g.shape[0])

    def __getitem__(self, k):
        return self.S[k].view(self.n_b, self.n_b).t(), self.g[k]  # .t(): col-major storage -> (row, col)

    def dense(self):
        n = len(self)
        return self.S.view(n, self.n_b, self.n_b).transpose(1, 2), self.g
