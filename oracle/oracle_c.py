"""ctypes wrapper of the C oracle twin (oracle/oracle_c.c).  TEST INFRASTRUCTURE ONLY -- see oracle.py."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle_c.so")
    src = os.path.join(_HERE, "oracle_c.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle_c.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        _LIB.ora_max_threads.restype = ctypes.c_int
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t))


def _plan_args(plan):
    ndofs = np.ascontiguousarray(plan.ndofs, dtype=np.int32)
    boff = np.ascontiguousarray(plan.block_offset, dtype=np.int64)
    perm = np.ascontiguousarray(plan.perm_fields, dtype=np.int32)
    return ndofs, boff, perm


def max_threads() -> int:
    return int(lib().ora_max_threads())


def condense(plan, A, b, nthreads=0):
    A = np.ascontiguousarray(A, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    ncells = A.shape[0]
    ndofs, boff, perm = _plan_args(plan)
    S = np.empty((ncells, plan.n_b * plan.n_b))
    g = np.empty((ncells, plan.n_b))
    info = np.empty(ncells, dtype=np.int32)
    rc = lib().ora_condense(ctypes.c_int(plan.nfields), _p(ndofs, ctypes.c_int32), _p(boff, ctypes.c_int64),
                            _p(perm, ctypes.c_int32), ctypes.c_int(len(plan.interior)), ctypes.c_int64(ncells),
                            _p(A, ctypes.c_double), ctypes.c_int64(plan.lenA), _p(b, ctypes.c_double),
                            ctypes.c_int64(plan.lenb), _p(S, ctypes.c_double), _p(g, ctypes.c_double),
                            _p(info, ctypes.c_int32), ctypes.c_int(nthreads))
    assert rc == 0
    return S, g, info


def backsub(plan, A, b, x, nthreads=0):
    A = np.ascontiguousarray(A, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    x = np.ascontiguousarray(x, dtype=np.float64)
    ncells = A.shape[0]
    ndofs, boff, perm = _plan_args(plan)
    u = np.empty((ncells, plan.n_i))
    info = np.empty(ncells, dtype=np.int32)
    rc = lib().ora_backsub(ctypes.c_int(plan.nfields), _p(ndofs, ctypes.c_int32), _p(boff, ctypes.c_int64),
                           _p(perm, ctypes.c_int32), ctypes.c_int(len(plan.interior)), ctypes.c_int64(ncells),
                           _p(A, ctypes.c_double), ctypes.c_int64(plan.lenA), _p(b, ctypes.c_double),
                           ctypes.c_int64(plan.lenb), _p(x, ctypes.c_double), _p(u, ctypes.c_double),
                           _p(info, ctypes.c_int32), ctypes.c_int(nthreads))
    assert rc == 0
    return u, info


def assemble_coo_csc(S, g, cell_ids, nfree):
    """Gridap's COO numeric loop + SparseArrays.sparse! (serial C restatement, oracle_c.c): S [ncells, n_b*n_b] column-major
    per cell, g [ncells, n_b], cell_ids [ncells, n_b] (1-based, <= 0 not assembled) -> (colptr, rowval, nzval, rhs),
    Int64 1-based like SparseMatrixCSC{Float64,Int64}."""
    S = np.ascontiguousarray(S, dtype=np.float64)
    g = np.ascontiguousarray(g, dtype=np.float64)
    ids = np.ascontiguousarray(cell_ids, dtype=np.int64)
    ncells, n_b = ids.shape
    cap = int(((ids > 0).sum(axis=1) ** 2).sum())
    colptr = np.empty(nfree + 1, dtype=np.int64)
    rowval = np.empty(max(cap, 1), dtype=np.int64)
    nzval = np.empty(max(cap, 1), dtype=np.float64)
    rhs = np.empty(nfree, dtype=np.float64)
    L = lib()
    L.ora_assemble_coo_csc.restype = ctypes.c_int64
    L.ora_assemble_coo_csc.argtypes = [ctypes.c_int64, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                       ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                       ctypes.c_int64]
    nnz = L.ora_assemble_coo_csc(ncells, n_b, ids.ctypes.data, S.ctypes.data, g.ctypes.data, nfree, colptr.ctypes.data,
                                 rowval.ctypes.data, nzval.ctypes.data, rhs.ctypes.data, cap)
    if nnz < 0:
        raise RuntimeError(f"ora_assemble_coo_csc failed: {nnz}")
    return colptr, rowval[:nnz].copy(), nzval[:nnz].copy(), rhs
