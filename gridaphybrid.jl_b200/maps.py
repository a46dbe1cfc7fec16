"""Host-side mirror of the reference's Map types for the hot path (same names, same argument meaning).

Reference (file:line relative to /root/reference):
  StaticCondensationMap ........... src/StaticCondensationMap.jl:3-34 (struct+checks), :152-196 (evaluate!)
  BackwardStaticCondensationMap ... src/BackwardStaticCondensationMap.jl:6-19, :61-112
  Scalar2ArrayBlockMap ............ src/Scalar2ArrayBlockMap.jl:1-69
  RestrictArrayBlockMap ........... src/RestrictArrayBlockMap.jl:1-32
  SumFacetsMap .................... src/SumFacetsMap.jl:1-30
  lazy_map sites .................. src/HybridAffineFEOperators.jl:338, :117-118, :147, :353

Field ids are 1-based as on the Julia side.  `lazy_map(k, cells)` is the array-level entry: one C-ABI
call for the whole cell array.  `evaluate(cache, ...)` keeps the per-cell contract (a batch of one).
All arithmetic runs in libgridaphybrid_b200.so on the GPU; there is no CPU path here.
"""
from __future__ import annotations

import numpy as np
import torch

from .blocks import ArrayBlock, CondensedCells, PackedCells, block_sizes
from .context import Context

_default_ctx = None


def default_context() -> Context:
    global _default_ctx
    if _default_ctx is None:
        dev = torch.cuda.current_device() if torch.cuda.is_available() else 0
        _default_ctx = Context(dev)  # raises without a CUDA device: no CPU fallback
    return _default_ctx


def set_default_context(ctx):
    global _default_ctx
    _default_ctx = ctx


def _check_preconditions(interior_fields, boundary_fields) -> bool:
    """src/StaticCondensationMap.jl:16-34."""
    nf = len(interior_fields) + len(boundary_fields)
    if not all(1 <= f <= nf for f in list(interior_fields) + list(boundary_fields)):
        return False
    seen = set()
    for f in list(interior_fields) + list(boundary_fields):
        if f in seen:
            return False
        seen.add(f)
    return len(seen) == nf


class StaticCondensationMap:
    """`(A_K, b_K) -> (S_K, g_K)`: S = A22 - A21 A11^-1 A12, g = b2 - A21 A11^-1 b1."""

    def __init__(self, interior_fields, boundary_fields):
        self.interior_fields = [int(f) for f in interior_fields]
        self.boundary_fields = [int(f) for f in boundary_fields]
        assert _check_preconditions(self.interior_fields, self.boundary_fields), \
            "interior_fields and boundary_fields must be a disjoint cover of 1:nfields"

    def plan(self, cells: PackedCells, ctx: Context):
        key = (id(ctx), tuple(cells.ndofs), cells.touched.tobytes(), tuple(self.interior_fields), tuple(self.boundary_fields))
        cache = self.__dict__.setdefault("_plans", {})
        if key not in cache:
            cache[key] = ctx.plan_blocks(cells.ndofs, cells.touched, self.interior_fields, self.boundary_fields)
        return cache[key]

    # per-cell protocol ------------------------------------------------------------------------
    def return_cache(self, A: ArrayBlock, b: ArrayBlock):
        brs, bcs = block_sizes(A)
        assert brs == bcs  # src/StaticCondensationMap.jl:52
        return {"brs": brs}

    def evaluate(self, cache, A: ArrayBlock, b: ArrayBlock):
        """`evaluate!(cache,k,A,b)`: returns (S, g) as float64 numpy arrays (row, col) of ONE cell."""
        out = lazy_map(self, PackedCells.from_cells([(A, b)], device="cuda"))
        S, g = out[0]
        info = int(out.info[0])
        assert info == 0, f"getrf info={info}"  # Gridap.Helpers.@check info==0 (:180)
        return S.cpu().numpy().copy(), g.cpu().numpy().copy()


class BackwardStaticCondensationMap:
    """`(A_K, b_K, lambda_K) -> [u_K by interior field..., lambda_K by boundary field...]`."""

    def __init__(self, interior_fields, boundary_fields):
        self.static_condensation = StaticCondensationMap(interior_fields, boundary_fields)

    def return_cache(self, A, b, x):
        return self.static_condensation.return_cache(A, b)

    def evaluate(self, cache, A: ArrayBlock, b: ArrayBlock, x):
        """x: dense vector or VectorBlock (densified = concatenated first, :104-112)."""
        if isinstance(x, ArrayBlock):
            x = np.concatenate([np.asarray(v, dtype=np.float64) for v, t in zip(x.array, x.touched) if t])
        cells = PackedCells.from_cells([(A, b)], device="cuda")
        xk = torch.as_tensor(np.asarray(x, dtype=np.float64)).reshape(1, -1).cuda()
        blocks = lazy_map(self, cells, xk)
        return ArrayBlock([v[0].cpu().numpy() for v in blocks], np.ones(len(blocks), dtype=bool))


class Scalar2ArrayBlockMap:
    """Dense `(S, g)` -> field-blocked `(MatrixBlock, VectorBlock)` by block sizes `bs`
    (src/Scalar2ArrayBlockMap.jl:42-69).  In the packed layout this is pure slicing (views)."""

    def return_cache(self, A, b, bs):
        return None

    def evaluate(self, cache, A, b, bs):
        off = np.concatenate([[0], np.cumsum(bs)])
        nb = len(bs)
        Ab = [[A[..., off[i]:off[i + 1], off[j]:off[j + 1]] for j in range(nb)] for i in range(nb)]
        bb = [b[..., off[i]:off[i + 1]] for i in range(nb)]
        return ArrayBlock(Ab, np.ones((nb, nb), dtype=bool)), ArrayBlock(bb, np.ones(nb, dtype=bool))


class RestrictArrayBlockMap:
    """Pick `blocks` (1-based) out of a VectorBlock (src/RestrictArrayBlockMap.jl:22-32)."""

    def __init__(self, blocks):
        self.blocks = [int(k) for k in blocks]

    def return_cache(self, v):
        return None

    def evaluate(self, cache, v: ArrayBlock):
        arr = [v.array[k - 1] if v.touched[k - 1] else None for k in self.blocks]
        return ArrayBlock(arr, [bool(v.touched[k - 1]) for k in self.blocks])


class SumFacetsMap:
    """In-cell reduction over local facets, res = a[1]+...+a[nlfacets] (src/SumFacetsMap.jl:19-30).

    Batched form: `a` float64 [ncells, nlfacets, ...] -> [ncells, ...] (facet contributions already laid
    out on the cell block; for facet-field blocks the summands are disjoint, so the sum is placement)."""

    def return_cache(self, a):
        return None

    def evaluate(self, cache, a):
        if isinstance(a, ArrayBlock):
            assert a.touched.all()  # :24
            res = _add(a.array[0], a.array[1])
            for i in range(2, len(a.array)):
                res = _add(res, a.array[i])
            return res
        # batched: [ncells, nlfacets, ...] on the device -> ghb_sum_facets_f64
        ctx = default_context()
        a = a.contiguous()
        n, nlf = int(a.shape[0]), int(a.shape[1])
        out = torch.empty((n,) + tuple(a.shape[2:]), dtype=torch.float64, device=a.device)
        ctx.use_torch_stream()
        ctx.sum_facets(n, nlf, a[0, 0].numel(), a, out)
        return out


def _add(x, y):
    if isinstance(x, ArrayBlock):
        touched = x.touched | y.touched
        fx, fy = _flat(x), _flat(y)
        out = []
        for tx, ty, ax, ay in zip(x.touched.ravel(), y.touched.ravel(), fx, fy):
            out.append(_add(ax, ay) if (tx and ty) else (ax if tx else (ay if ty else None)))
        if touched.ndim == 2:
            c = touched.shape[1]
            out = [out[i * c:(i + 1) * c] for i in range(touched.shape[0])]
        return ArrayBlock(out, touched)
    return x + y


def _flat(a: ArrayBlock):
    return list(a.array) if a.touched.ndim == 1 else [e for row in a.array for e in row]


# ------------------------------------------------------------------------------------------------
# lazy_map: the array-level drop-in sites
# ------------------------------------------------------------------------------------------------


def lazy_map(k, t, *args, ctx: Context | None = None, keep_factors: bool = False):
    """`lazy_map(k, t, args...)` specialised on the Map type, evaluated eagerly on the GPU.

    StaticCondensationMap:          t: PackedCells                  -> CondensedCells
    BackwardStaticCondensationMap:  t: PackedCells, lhk [ncells,n_b] -> list of per-field tensors
                                    [u fields..., lambda fields...]  (block positions 1..|I|, |I|+1..)
    Scalar2ArrayBlockMap:           t: CondensedCells, bs            -> (MatrixBlock, VectorBlock) of views
    RestrictArrayBlockMap:          t: list of per-field tensors     -> selected list
    """
    ctx = ctx or default_context()
    if isinstance(k, StaticCondensationMap):
        plan = k.plan(t, ctx)
        dev = torch.device("cuda", ctx.device)
        n = len(t)
        S = torch.empty((n, plan.n_b * plan.n_b), dtype=torch.float64, device=dev)
        g = torch.empty((n, plan.n_b), dtype=torch.float64, device=dev)
        info = torch.empty((n,), dtype=torch.int32, device=dev)
        ctx.use_torch_stream()
        if hasattr(t, "family"):                 # AffineCells: records formed in the loader of the condensation kernel
            t.family.condense(ctx, plan, t.coef, S, g, info, keep_factors=keep_factors)
            return CondensedCells(S, g, info, plan.n_b, plan)
        ctx.condense(plan, n, t.A, t.b, S, g, info, keep_factors=keep_factors)
        return CondensedCells(S, g, info, plan.n_b, plan)
    if isinstance(k, BackwardStaticCondensationMap):
        (lhk,) = args
        sc = k.static_condensation
        plan = sc.plan(t, ctx)
        dev = torch.device("cuda", ctx.device)
        n = len(t)
        if isinstance(lhk, (list, tuple)):  # VectorBlock lambda: densify = concatenate (:104-112)
            lhk = torch.cat([torch.as_tensor(v) for v in lhk], dim=1)
        lhk = torch.as_tensor(lhk, dtype=torch.float64).to(dev).contiguous()
        assert lhk.shape == (n, plan.n_b)
        # cell-wise values are addressed through synthetic ids 1..n*n_b into the flattened lhk
        ids = torch.arange(1, n * plan.n_b + 1, dtype=torch.int64, device=dev).view(n, plan.n_b)
        u = torch.empty((n, plan.n_i), dtype=torch.float64, device=dev)
        info = torch.empty((n,), dtype=torch.int32, device=dev)
        ctx.use_torch_stream()
        ctx.backsub(plan, n, t.A, t.b, lhk.view(-1), None, ids, u, info)
        k.last_info = info
        out, o = [], 0
        for f in sc.interior_fields:
            s = t.ndofs[f - 1]
            out.append(u[:, o:o + s])
            o += s
        o = 0
        for f in sc.boundary_fields:
            s = t.ndofs[f - 1]
            out.append(lhk[:, o:o + s])
            o += s
        return out
    if isinstance(k, Scalar2ArrayBlockMap):
        (bs,) = args
        S, g = t.dense()
        return k.evaluate(None, S, g, bs)
    if isinstance(k, RestrictArrayBlockMap):
        return [t[b - 1] for b in k.blocks]
    raise TypeError(f"lazy_map: unsupported map {type(k)}")


def _touched_blocks(X: ArrayBlock):
    """`findall(X.touched)` in Julia's (column-major) order, 0-based index tuples"""
    t = X.touched
    if t.ndim == 1:
        return [(i,) for i in range(t.shape[0]) if t[i]]
    return [(i, j) for j in range(t.shape[1]) for i in range(t.shape[0]) if t[i, j]]


def _block(X: ArrayBlock, idx):
    return X.array[idx[0]] if len(idx) == 1 else X.array[idx[0]][idx[1]]


def compute_bulk_to_skeleton_l2_projection_dofs(A, B, ctx: Context | None = None, info=None):
    """`lazy_map(compute_bulk_to_skeleton_l2_projection_dofs, A_array, B_array)` of test/P_m.jl:4-23, evaluated on the
    batch (src/GridapAPIExtensions.jl:453-500: `A\\B` per (cell, local facet)).  A [nbatch, n, n] facet mass matrices,
    B [nbatch, n, m] bulk-basis moments or [nbatch, n] (FE function) -> X like B (device tensors).

    Block arguments (`ArrayBlock`s whose entries carry the batch as their leading dimension) follow the reference's
    overloads, src/GridapAPIExtensions.jl:547-760:
      * (VectorBlock, VectorBlock), same touched mask: the map applied to every touched entry (`:551-590`; the entries
        may be blocks themselves) -> VectorBlock;
      * (MatrixBlock, MatrixBlock), one touched block each, in the same block row (`:617-663`): `A[.,b1] \\ B[.,b2]`
        placed at `[1, b2]` of a `1 x nb` MatrixBlock;
      * (MatrixBlock, VectorBlock of vectors), one touched block each (`:665-697`): the plain array `A_blk \\ B_blk`;
      * (MatrixBlock, VectorBlock of matrices), one touched block each (`:699-742`): a VectorBlock of length 1."""
    if isinstance(A, ArrayBlock) and isinstance(B, ArrayBlock):
        f = compute_bulk_to_skeleton_l2_projection_dofs
        if A.touched.ndim == 1 and B.touched.ndim == 1:
            assert A.touched.shape == B.touched.shape and np.array_equal(A.touched, B.touched)
            r = [f(A.array[i], B.array[i], ctx) if A.touched[i] else None for i in range(len(A.array))]
            return ArrayBlock(r, A.touched.copy())
        nA, nB = _touched_blocks(A), _touched_blocks(B)
        assert len(nA) == len(nB) == 1, "exactly one touched block in A and in B"
        ai, bi = _block(A, nA[0]), _block(B, nB[0])
        if A.touched.ndim == 2 and B.touched.ndim == 2:
            assert A.touched.shape == B.touched.shape and nA[0][0] == nB[0][0]
            tb, nb = nB[0][1], A.touched.shape[0]
            touched = np.zeros((1, nb), dtype=bool)
            touched[0, tb] = True
            row = [None] * nb
            row[tb] = f(ai, bi, ctx)
            return ArrayBlock([row], touched)
        assert A.touched.ndim == 2 and B.touched.ndim == 1 and A.touched.shape[1] == B.touched.shape[0]
        x = f(ai, bi, ctx)
        if torch.as_tensor(bi).dim() == 2:          # blocks of B are vectors (one right-hand side): plain array
            return x
        assert nA[0][0] == nB[0][0] and nA[0][1] == nB[0][0] and A.touched.shape[0] == A.touched.shape[1]
        return ArrayBlock([x], np.ones(1, dtype=bool))
    ctx = ctx or default_context()
    dev = torch.device("cuda", ctx.device)
    A = torch.as_tensor(A, dtype=torch.float64).to(dev)
    B = torch.as_tensor(B, dtype=torch.float64).to(dev)
    vec = B.dim() == 2
    nb, n = int(A.shape[0]), int(A.shape[1])
    assert A.shape == (nb, n, n) and B.shape[:2] == (nb, n)
    m = 1 if vec else int(B.shape[2])
    Ac = A.transpose(1, 2).contiguous()                       # column-major per system
    Bc = B.reshape(nb, n, m).transpose(1, 2).contiguous()
    X = torch.empty_like(Bc)
    ctx.use_torch_stream()
    ctx.l2_projection_dofs(nb, n, m, Ac, Bc, X, info)
    X = X.transpose(1, 2)
    return X[:, :, 0].contiguous() if vec else X.contiguous()
