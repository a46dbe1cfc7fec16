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
