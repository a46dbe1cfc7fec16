"""CPU tests of the drop-in boundary: the C-ABI library builds, loads and exports every symbol that
include/ghb.h declares; without a GPU the product fails loudly (no CPU fallback)."""
import ctypes
import os
import re

import pytest

import gridaphybrid_b200 as gh
from gridaphybrid_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "ghb.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(ghb_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_exported():
    L = gh.lib()
    declared = _declared_symbols()
    assert len(declared) >= 15
    for s in declared:
        assert hasattr(L, s), f"{s} declared in include/ghb.h but not exported"
    assert sorted(_lib.SYMBOLS) == declared


def test_library_is_in_tree_and_sm100a():
    assert os.path.exists(_lib.SO_PATH) and _lib.SO_PATH.startswith(ROOT)
    assert "arch=compute_100a,code=sm_100a" in _lib.NVCC_FLAGS


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(gh.GhbError) as e:
        gh.Context(0)
    assert e.value.code == _lib.GHB_ENODEVICE
    h = ctypes.c_void_p()
    assert gh.lib().ghb_create(0, ctypes.byref(h)) == _lib.GHB_ENODEVICE


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "gridaphybrid.jl_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "liboracle" not in src, f


def test_map_preconditions():
    # src/StaticCondensationMap.jl:16-34
    gh.StaticCondensationMap([1, 2], [3])
    with pytest.raises(AssertionError):
        gh.StaticCondensationMap([1, 2], [2])
    with pytest.raises(AssertionError):
        gh.StaticCondensationMap([1, 4], [3])


def test_packed_layout():
    import numpy as np
    off, n = gh.PackedCells.layout([2, 1, 3], np.array([[1, 1, 1], [1, 0, 0], [1, 0, 1]], bool))
    assert n == 4 + 2 + 6 + 2 + 6 + 9
    assert off.tolist() == [[0, 12, 14], [4, -1, -1], [6, -1, 20]]
