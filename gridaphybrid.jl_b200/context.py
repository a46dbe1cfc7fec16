"""`Context`: thin object wrapper over the C ABI (include/ghb.h).  One Context per process / GPU.

Arrays may be torch tensors (CUDA or CPU) or numpy arrays (host); the pointer is handed to the C
library unchanged -- the library itself decides host vs device (cudaPointerGetAttributes).
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib
from ._lib import GhbError

try:  # torch is plumbing (device memory + streams); numpy-only callers work with host pointers
    import torch
except Exception:  # pragma: no cover
    torch = None


def _len(x):
    return 0 if x is None else int(x.numel() if hasattr(x, "numel") else x.size)


def _ptr(x, dtype=None, count=None, name="array"):
    if x is None:
        return None
    if torch is not None and isinstance(x, torch.Tensor):
        if not x.is_contiguous():
            raise ValueError(f"{name}: tensor must be contiguous")
        if dtype is not None:
            want = {np.float64: torch.float64, np.int64: torch.int64, np.int32: torch.int32, np.uint8: torch.uint8}[dtype]
            if x.dtype != want:
                raise TypeError(f"{name}: expected {want}, got {x.dtype}")
        if count is not None and x.numel() < count:
            raise ValueError(f"{name}: needs {count} elements, has {x.numel()}")
        return ctypes.c_void_p(x.data_ptr())
    if isinstance(x, np.ndarray):
        if not x.flags["C_CONTIGUOUS"]:
            raise ValueError(f"{name}: array must be C-contiguous")
        if dtype is not None and x.dtype != np.dtype(dtype):
            raise TypeError(f"{name}: expected {np.dtype(dtype)}, got {x.dtype}")
        if count is not None and x.size < count:
            raise ValueError(f"{name}: needs {count} elements, has {x.size}")
        return ctypes.c_void_p(x.ctypes.data)
    raise TypeError(f"{name}: expected torch.Tensor or numpy.ndarray, got {type(x)}")


class BlockPlan:
    """Handle of a block plan living inside a Context (ghb_plan_blocks)."""

    def __init__(self, ctx, plan_id, ndofs, touched, interior, boundary, n_i, n_b, lenA, lenb):
        self.ctx, self.id = ctx, plan_id
        self.ndofs, self.touched = list(ndofs), np.array(touched, dtype=bool)
        self.interior, self.boundary = list(interior), list(boundary)
        self.n_i, self.n_b, self.lenA, self.lenb = n_i, n_b, lenA, lenb
        self.n = n_i + n_b

    @property
    def kernel_name(self) -> str:
        return _lib.lib().ghb_plan_kernel_name(self.ctx._h, self.id).decode()


class Context:
    def __init__(self, device: int = 0):
        self._L = _lib.lib()
        h = ctypes.c_void_p()
        rc = self._L.ghb_create(int(device), ctypes.byref(h))
        if rc != 0:
            if rc == _lib.GHB_ENODEVICE:
                raise GhbError(rc, "no CUDA device: libgridaphybrid_b200 has no CPU fallback")
            raise GhbError(rc, "ghb_create failed")
        self._h = h
        self.device = int(device)
        self._asm_shapes = {}
        if torch is not None and torch.cuda.is_available():
            self.use_torch_stream()   # stream-ordered with the caller's torch work from the start

    def close(self):
        if getattr(self, "_h", None):
            self._L.ghb_destroy(self._h)
            self._h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise GhbError(rc, self._L.ghb_last_error(self._h).decode())

    # ---- plumbing -------------------------------------------------------------------------
    def use_torch_stream(self):
        """Launch on torch's current CUDA stream (so torch.cuda.Event timing sees the kernels)."""
        s = torch.cuda.current_stream(self.device).cuda_stream
        self._check(self._L.ghb_set_stream(self._h, ctypes.c_void_p(s)))

    def synchronize(self):
        self._check(self._L.ghb_synchronize(self._h))

    @property
    def launch_count(self) -> int:
        return int(self._L.ghb_launch_count(self._h))

    def host_register(self, arr):
        """page-lock a host array in place (numpy array or CPU tensor); pair with host_unregister before it is freed"""
        a = arr.numpy() if hasattr(arr, "numpy") and not isinstance(arr, np.ndarray) else arr
        self._check(self._L.ghb_host_register(self._h, ctypes.c_void_p(a.ctypes.data), int(a.nbytes)))

    def host_unregister(self, arr):
        a = arr.numpy() if hasattr(arr, "numpy") and not isinstance(arr, np.ndarray) else arr
        self._check(self._L.ghb_host_unregister(self._h, ctypes.c_void_p(a.ctypes.data)))

    def set_option(self, name: str, value: int):
        """debugging / A-B knobs (ghb_set_option): "cw", "dmma_ll", "force_generic", "factors_generic",
        "max_ctas_per_sm", "stream_chunk_bytes", ...; kernel-choice knobs act on plans created afterwards"""
        self._check(self._L.ghb_set_option(self._h, name.encode(), int(value)))

    # ---- plans ----------------------------------------------------------------------------
    def plan_blocks(self, ndofs, touched, interior_fields, boundary_fields) -> BlockPlan:
        ndofs_a = np.ascontiguousarray(ndofs, dtype=np.int32)
        nf = len(ndofs_a)
        t = np.asarray(touched, dtype=bool).reshape(nf, nf)
        t_cm = np.ascontiguousarray(t.T.astype(np.uint8))  # column-major for the ABI
        ia = np.ascontiguousarray(interior_fields, dtype=np.int32)
        ba = np.ascontiguousarray(boundary_fields, dtype=np.int32)
        pid = ctypes.c_int(-1)
        self._check(self._L.ghb_plan_blocks(self._h, nf, _ptr(ndofs_a), _ptr(t_cm), len(ia),
                                            _ptr(ia) if len(ia) else None, len(ba), _ptr(ba) if len(ba) else None,
                                            ctypes.byref(pid)))
        q = (ctypes.c_int64 * 4)()
        self._check(self._L.ghb_plan_query(self._h, pid.value, q))
        return BlockPlan(self, pid.value, ndofs_a.tolist(), t, ia.tolist(), ba.tolist(), int(q[0]), int(q[1]),
                         int(q[2]), int(q[3]))

    # ---- hot path ---------------------------------------------------------------------------
    def condense(self, plan: BlockPlan, ncells, A, b, S, g, info=None, keep_factors=False):
        self._check(self._L.ghb_condense_f64(
            self._h, plan.id, int(ncells), _ptr(A, np.float64, ncells * plan.lenA, "A"),
            _ptr(b, np.float64, ncells * plan.lenb, "b"), _ptr(S, np.float64, ncells * plan.n_b ** 2, "S"),
            _ptr(g, np.float64, ncells * plan.n_b, "g"), _ptr(info, np.int32, ncells, "info"), int(bool(keep_factors))))

    def restrict_facet_dofs(self, ncells, nlfacets, ndofs_f, cell_wise_facets, facet_data, out):
        self._check(self._L.ghb_restrict_facet_dofs_i64(
            self._h, int(ncells), int(nlfacets), int(ndofs_f), _ptr(cell_wise_facets, np.int64, ncells * nlfacets),
            _ptr(facet_data, np.int64), _ptr(out, np.int64, ncells * nlfacets * ndofs_f)))

    def sum_facets(self, ncells, nlfacets, length, a, out):
        self._check(self._L.ghb_sum_facets_f64(self._h, int(ncells), int(nlfacets), int(length),
                                               _ptr(a, np.float64, ncells * nlfacets * length, "in"),
                                               _ptr(out, np.float64, ncells * length, "out")))

    def l2_projection_dofs(self, nbatch, n, nrhs, A, B, X, info=None):
        """X = A \\ B for a batch of n x n systems with nrhs right-hand sides (column-major per system)."""
        self._check(self._L.ghb_l2_projection_dofs_f64(
            self._h, int(nbatch), int(n), int(nrhs), _ptr(A, np.float64, nbatch * n * n, "A"),
            _ptr(B, np.float64, nbatch * n * nrhs, "B"), _ptr(X, np.float64, nbatch * n * nrhs, "X"),
            _ptr(info, np.int32, nbatch, "info")))

    def expand_records(self, plan: BlockPlan, ncells, ntab, TA, Tb, coef, A, b):
        """A_K = sum_t coef[K][t] TA[t], b_K likewise (records of an affine family, generated on the device)."""
        self._check(self._L.ghb_expand_records_f64(
            self._h, plan.id, int(ncells), int(ntab), _ptr(TA, np.float64, ntab * plan.lenA, "TA"),
            _ptr(Tb, np.float64, ntab * plan.lenb, "Tb"), _ptr(coef, np.float64, ncells * ntab, "coef"),
            _ptr(A, np.float64, ncells * plan.lenA, "A"), _ptr(b, np.float64, ncells * plan.lenb, "b")))

    def condense_affine(self, plan: BlockPlan, ncells, ntab, TA, Tb, coef, S, g, info=None, keep_factors=False):
        """S_K, g_K of an affine family without materialising the records (generated in the loader of the condensation
        kernel); bit-identical to expand_records + condense."""
        self._check(self._L.ghb_condense_affine_f64(
            self._h, plan.id, int(ncells), int(ntab), _ptr(TA, np.float64, ntab * plan.lenA, "TA"),
            _ptr(Tb, np.float64, ntab * plan.lenb, "Tb"), _ptr(coef, np.float64, ncells * ntab, "coef"),
            _ptr(S, np.float64, ncells * plan.n_b ** 2, "S"), _ptr(g, np.float64, ncells * plan.n_b, "g"),
            _ptr(info, np.int32, ncells, "info"), int(bool(keep_factors))))

    def condense_assemble_affine(self, plan: BlockPlan, ncells, ntab, TA, Tb, coef, dirichlet_vals, nzval, rhs, info=None):
        """coefficients -> CSC values + rhs in one call (selected symbolic pattern)."""
        nrows, nnz = self._asm_shape
        self._check(self._L.ghb_condense_assemble_affine_f64(
            self._h, plan.id, int(ncells), int(ntab), _ptr(TA, np.float64, ntab * plan.lenA, "TA"),
            _ptr(Tb, np.float64, ntab * plan.lenb, "Tb"), _ptr(coef, np.float64, ncells * ntab, "coef"),
            _ptr(dirichlet_vals, np.float64), _len(dirichlet_vals), _ptr(nzval, np.float64, nnz, "nzval"),
            _ptr(rhs, np.float64, nrows, "rhs"), _ptr(info, np.int32, ncells, "info")))

    def backsub_affine(self, plan: BlockPlan, ncells, ntab, TA, Tb, coef, lambda_free, lambda_dirichlet, cell_ids, u, info=None):
        """backward map of an affine family with the records formed in the loader (ghb_backsub_affine_f64)"""
        self._check(self._L.ghb_backsub_affine_f64(
            self._h, plan.id, int(ncells), int(ntab), _ptr(TA, np.float64, ntab * plan.lenA, "TA"),
            _ptr(Tb, np.float64, ntab * plan.lenb, "Tb"), _ptr(coef, np.float64, ncells * ntab, "coef"),
            _ptr(lambda_free, np.float64), _len(lambda_free), _ptr(lambda_dirichlet, np.float64), _len(lambda_dirichlet),
            _ptr(cell_ids, np.int64, ncells * plan.n_b, "cell_ids"), _ptr(u, np.float64, ncells * plan.n_i, "u"),
            _ptr(info, np.int32, ncells, "info")))

    def assemble_symbolic(self, ncells, n_b, cell_ids, nrows) -> int:
        nnz = ctypes.c_int64(0)
        self._check(self._L.ghb_assemble_symbolic(self._h, int(ncells), int(n_b),
                                                  _ptr(cell_ids, np.int64, ncells * n_b, "cell_ids"), int(nrows),
                                                  ctypes.byref(nnz)))
        self._asm_shape = (int(nrows), int(nnz.value))
        self._asm_shapes[self.assemble_current()] = self._asm_shape
        return int(nnz.value)

    def assemble_current(self) -> int:
        """handle of the selected symbolic pattern (every symbolic call creates and selects a new one)"""
        return int(self._L.ghb_assemble_current(self._h))

    def assemble_select(self, pattern_id: int):
        self._check(self._L.ghb_assemble_select(self._h, int(pattern_id)))
        self._asm_shape = self._asm_shapes[int(pattern_id)]

    def assemble_release(self, pattern_id: int):
        self._check(self._L.ghb_assemble_release(self._h, int(pattern_id)))
        self._asm_shapes.pop(int(pattern_id), None)

    def assemble_pattern(self, colptr, rowval):
        nrows, nnz = self._asm_shape
        self._check(self._L.ghb_assemble_pattern(self._h, _ptr(colptr, np.int64, nrows + 1, "colptr"),
                                                 _ptr(rowval, np.int64, nnz, "rowval")))

    def assemble_numeric(self, S, g, dirichlet_vals, nzval, rhs):
        nrows, nnz = self._asm_shape
        self._check(self._L.ghb_assemble_numeric_f64(self._h, _ptr(S, np.float64), _ptr(g, np.float64),
                                                     _ptr(dirichlet_vals, np.float64), _len(dirichlet_vals),
                                                     _ptr(nzval, np.float64, nnz, "nzval"),
                                                     _ptr(rhs, np.float64, nrows, "rhs")))

    def assemble_numeric_csr(self, S, g, dirichlet_vals, nzval, rhs):
        """CSR values of the skeleton matrix (pattern = the CSC pattern, structurally symmetric); transposes the
        cell blocks of the device array `S` in place."""
        nrows, nnz = self._asm_shape
        self._check(self._L.ghb_assemble_numeric_csr_f64(self._h, _ptr(S, np.float64), _ptr(g, np.float64),
                                                         _ptr(dirichlet_vals, np.float64), _len(dirichlet_vals),
                                                         _ptr(nzval, np.float64, nnz, "nzval"),
                                                         _ptr(rhs, np.float64, nrows, "rhs")))

    def assemble_symbolic_slab(self, ncells_local, nghost, ghost_ncols, n_b, cell_ids, nrows_global, col_begin,
                               col_end) -> int:
        nnz = ctypes.c_int64(0)
        self._check(self._L.ghb_assemble_symbolic_slab(
            self._h, int(ncells_local), int(nghost), int(ghost_ncols), int(n_b),
            _ptr(cell_ids, np.int64, (ncells_local + nghost) * n_b, "cell_ids"), int(nrows_global), int(col_begin),
            int(col_end), ctypes.byref(nnz)))
        self._asm_shape = (int(col_end - col_begin), int(nnz.value))
        self._asm_shapes[self.assemble_current()] = self._asm_shape
        return int(nnz.value)

    def pack_cut_plane(self, ncut, n_b, ncols, S, g, cell_ids, dirichlet_vals, out):
        self._check(self._L.ghb_pack_cut_plane_f64(self._h, int(ncut), int(n_b), int(ncols), _ptr(S, np.float64),
                                                   _ptr(g, np.float64), _ptr(cell_ids, np.int64),
                                                   _ptr(dirichlet_vals, np.float64),
                                                   _ptr(out, np.float64, ncut * (n_b * ncols + ncols), "out")))

    def assemble_numeric_slab(self, S, g, ghost, dirichlet_vals, nzval, rhs):
        nrows, nnz = self._asm_shape
        self._check(self._L.ghb_assemble_numeric_slab_f64(self._h, _ptr(S, np.float64), _ptr(g, np.float64),
                                                          _ptr(ghost, np.float64), _ptr(dirichlet_vals, np.float64),
                                                          _ptr(nzval, np.float64, nnz, "nzval"),
                                                          _ptr(rhs, np.float64, nrows, "rhs")))

    def condense_scatter_slab(self, plan, ncells, A, b, S, g, info, nzval, keep_cut, zero_nzval=True):
        """fused condensation + scatter of the local cells of a slab (device arrays); S only where it is needed later"""
        nrows, nnz = self._asm_shape
        self._check(self._L.ghb_condense_scatter_slab_f64(
            self._h, plan.id, int(ncells), _ptr(A, np.float64, ncells * plan.lenA, "A"), _ptr(b, np.float64, ncells * plan.lenb, "b"),
            _ptr(S, np.float64, ncells * plan.n_b ** 2, "S"), _ptr(g, np.float64, ncells * plan.n_b, "g"),
            _ptr(info, np.int32, ncells, "info"), _ptr(nzval, np.float64, nnz, "nzval"), int(keep_cut), int(bool(zero_nzval))))

    def condense_scatter_slab_affine(self, plan, ncells, ntab, TA, Tb, coef, S, g, info, nzval, keep_cut, zero_nzval=True):
        """the same with the records of an affine family formed in the loader (device arrays)"""
        nrows, nnz = self._asm_shape
        self._check(self._L.ghb_condense_scatter_slab_affine_f64(
            self._h, plan.id, int(ncells), int(ntab), _ptr(TA, np.float64, ntab * plan.lenA, "TA"),
            _ptr(Tb, np.float64, ntab * plan.lenb, "Tb"), _ptr(coef, np.float64, ncells * ntab, "coef"),
            _ptr(S, np.float64, ncells * plan.n_b ** 2, "S"), _ptr(g, np.float64, ncells * plan.n_b, "g"),
            _ptr(info, np.int32, ncells, "info"), _ptr(nzval, np.float64, nnz, "nzval"), int(keep_cut), int(bool(zero_nzval))))

    def assemble_finish_slab(self, S, g, ghost, dirichlet_vals, nzval, rhs):
        nrows, nnz = self._asm_shape
        self._check(self._L.ghb_assemble_finish_slab_f64(self._h, _ptr(S, np.float64), _ptr(g, np.float64),
                                                         _ptr(ghost, np.float64), _ptr(dirichlet_vals, np.float64),
                                                         _ptr(nzval, np.float64, nnz, "nzval"),
                                                         _ptr(rhs, np.float64, nrows, "rhs")))

    def condense_assemble(self, plan, ncells, A, b, dirichlet_vals, nzval, rhs, info=None):
        nrows, nnz = self._asm_shape
        self._check(self._L.ghb_condense_assemble_f64(
            self._h, plan.id, int(ncells), _ptr(A, np.float64, ncells * plan.lenA, "A"),
            _ptr(b, np.float64, ncells * plan.lenb, "b"), _ptr(dirichlet_vals, np.float64), _len(dirichlet_vals),
            _ptr(nzval, np.float64, nnz, "nzval"), _ptr(rhs, np.float64, nrows, "rhs"),
            _ptr(info, np.int32, ncells, "info")))

    def backsub(self, plan, ncells, A, b, lambda_free, lambda_dirichlet, cell_ids, u, info=None):
        self._check(self._L.ghb_backsub_f64(
            self._h, plan.id, int(ncells), _ptr(A, np.float64, ncells * plan.lenA, "A"),
            _ptr(b, np.float64, ncells * plan.lenb, "b"), _ptr(lambda_free, np.float64), _len(lambda_free),
            _ptr(lambda_dirichlet, np.float64), _len(lambda_dirichlet),
            _ptr(cell_ids, np.int64, ncells * plan.n_b, "cell_ids"),
            _ptr(u, np.float64, ncells * plan.n_i, "u"), _ptr(info, np.int32, ncells, "info")))

    def scatter_free_dof_values(self, plan, ncells, u, lambda_free, x):
        nl = 0 if lambda_free is None else (lambda_free.numel() if hasattr(lambda_free, "numel") else lambda_free.size)
        self._check(self._L.ghb_scatter_free_dof_values(
            self._h, plan.id, int(ncells), _ptr(u, np.float64, ncells * plan.n_i, "u"),
            _ptr(lambda_free, np.float64), int(nl), _ptr(x, np.float64, ncells * plan.n_i + nl, "x")))

    @property
    def factors_generation(self) -> int:
        return int(self._L.ghb_factors_generation(self._h))

    # ---- multi-GPU exchanges behind the C ABI (csrc/comm.cu) -------------------------------------
    def comm_unique_id(self) -> bytes:
        buf = ctypes.create_string_buffer(128)
        self._check(self._L.ghb_comm_unique_id(self._h, buf))
        return buf.raw

    def comm_init(self, nranks: int, rank: int, unique_id: bytes):
        assert len(unique_id) == 128
        self._check(self._L.ghb_comm_init(self._h, int(nranks), int(rank), ctypes.create_string_buffer(unique_id, 128)))
        self.comm_rank, self.comm_size = int(rank), int(nranks)

    def comm_init_from_torch(self, group=None):
        """ship rank 0's ncclUniqueId through torch.distributed (plumbing only) and create the communicator"""
        import torch.distributed as dist
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        box = [self.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        self.comm_init(world, rank, box[0])

    def comm_destroy(self):
        self._check(self._L.ghb_comm_destroy(self._h))

    def exchange_cut_plane(self, send_down, recv_from_up):
        self._check(self._L.ghb_exchange_cut_plane_f64(self._h, _ptr(send_down, np.float64), _len(send_down),
                                                       _ptr(recv_from_up, np.float64), _len(recv_from_up)))

    def allgather_lambda(self, owned, counts, out):
        c = np.ascontiguousarray(counts, dtype=np.int64)
        self._check(self._L.ghb_allgather_lambda_f64(self._h, _ptr(owned, np.float64), _ptr(c),
                                                     _ptr(out, np.float64, int(c.sum()), "all")))

    # ---- synthetic workload -------------------------------------------------------------------
    def synth_fill(self, plan, cell_start, ncells, A, b, seed=20261017):
        self._check(self._L.ghb_synth_fill_f64(self._h, plan.id, int(cell_start), int(ncells), int(seed),
                                               _ptr(A, np.float64, ncells * plan.lenA, "A"),
                                               _ptr(b, np.float64, ncells * plan.lenb, "b")))

    def cartesian_cell_wise_facets(self, dims, cell_start, ncells, out):
        d = np.ascontiguousarray(dims, dtype=np.int64)
        self._check(self._L.ghb_cartesian_cell_wise_facets(self._h, len(d), _ptr(d), int(cell_start), int(ncells),
                                                           _ptr(out, np.int64, ncells * 2 * len(d), "cell_wise_facets")))
