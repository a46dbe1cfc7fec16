"""Build + ctypes binding of libgridaphybrid_b200.so (the C ABI declared in include/ghb.h).

The library is the product; this module only loads it.  There is no CPU fallback: if the shared
library is missing it is compiled with nvcc for sm_100a, and if no CUDA device is present
`Context()` raises (ghb_create -> GHB_ENODEVICE).
"""
from __future__ import annotations

import ctypes
import glob
import os
import subprocess

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG)
_CSRC = os.path.join(_PKG, "csrc")
SO_PATH = os.path.join(_PKG, "libgridaphybrid_b200.so")
if os.environ.get("GHB_LIB_PATH"):      # A/B builds of the same sources with other -D knobs (tools/ab_variants.sh)
    SO_PATH = os.path.abspath(os.environ["GHB_LIB_PATH"])

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared", "-ldl"]

# every symbol include/ghb.h declares (tests check the library exports each one)
SYMBOLS = [
    "ghb_create", "ghb_destroy", "ghb_last_error", "ghb_set_stream", "ghb_synchronize", "ghb_launch_count",
    "ghb_set_option", "ghb_device_alloc", "ghb_device_free", "ghb_copy", "ghb_host_register", "ghb_host_unregister", "ghb_factors_generation",
    "ghb_assemble_current", "ghb_assemble_select", "ghb_assemble_release",
    "ghb_condense_scatter_slab_f64", "ghb_condense_scatter_slab_affine_f64", "ghb_assemble_finish_slab_f64",
    "ghb_comm_unique_id", "ghb_comm_init", "ghb_comm_destroy", "ghb_exchange_cut_plane_f64", "ghb_allgather_lambda_f64",
    "ghb_plan_kernel_name", "ghb_plan_blocks", "ghb_plan_query", "ghb_condense_f64",
    "ghb_restrict_facet_dofs_i64", "ghb_sum_facets_f64", "ghb_expand_records_f64", "ghb_condense_affine_f64", "ghb_condense_assemble_affine_f64", "ghb_backsub_affine_f64", "ghb_l2_projection_dofs_f64", "ghb_assemble_symbolic", "ghb_assemble_pattern", "ghb_assemble_numeric_f64", "ghb_assemble_numeric_csr_f64",
    "ghb_assemble_symbolic_slab", "ghb_pack_cut_plane_f64", "ghb_assemble_numeric_slab_f64",
    "ghb_condense_assemble_f64", "ghb_backsub_f64", "ghb_scatter_free_dof_values", "ghb_synth_fill_f64",
    "ghb_cartesian_cell_wise_facets",
]

GHB_OK, GHB_EINVAL, GHB_ECUDA, GHB_ENODEVICE, GHB_ENOMEM, GHB_EUNSUPPORTED, GHB_ESTATE = 0, -1, -2, -3, -4, -5, -6
_CODES = {-1: "GHB_EINVAL", -2: "GHB_ECUDA", -3: "GHB_ENODEVICE", -4: "GHB_ENOMEM", -5: "GHB_EUNSUPPORTED",
          -6: "GHB_ESTATE"}


class GhbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"{_CODES.get(code, code)}: {msg}")
        self.code = code


def sources():
    return sorted(glob.glob(os.path.join(_CSRC, "*.cu")))


def _source_hash() -> str:
    """content hash of every source the library is built from plus the build flags (mtimes do not survive a copy of
    the tree to another box; the hash does)"""
    import hashlib
    h = hashlib.sha256((" ".join(NVCC_FLAGS) + "|" + os.environ.get("GHB_NVCC_EXTRA", "")).encode())
    for d in sources() + sorted(glob.glob(os.path.join(_CSRC, "*.cuh"))) + [os.path.join(_ROOT, "include", "ghb.h")]:
        with open(d, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def needs_build() -> bool:
    """True if the library is missing or was built from other sources / flags (hash kept next to the .so)."""
    if not os.path.exists(SO_PATH):
        return True
    try:
        with open(SO_PATH + ".srchash") as f:
            return f.read().strip() != _source_hash()
    except OSError:
        return True


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every .cu under csrc/ for sm_100a into one in-tree shared library.  The translation units are compiled in
    parallel and their objects cached by content hash (sources + headers + flags) under csrc/_obj/, so an A/B build
    that changes one knob of one kernel recompiles one file."""
    if force or needs_build():
        import hashlib
        from concurrent.futures import ThreadPoolExecutor
        nvcc = os.environ.get("NVCC", "nvcc")
        extra = os.environ.get("GHB_NVCC_EXTRA", "").split()
        cflags = [f for f in NVCC_FLAGS if f not in ("-shared", "-ldl")] + extra
        hdr = hashlib.sha256()
        for d in sorted(glob.glob(os.path.join(_CSRC, "*.cuh"))) + [os.path.join(_ROOT, "include", "ghb.h")]:
            with open(d, "rb") as f:
                hdr.update(f.read())
        objdir = os.path.join(_CSRC, "_obj")
        os.makedirs(objdir, exist_ok=True)

        def compile_one(src):
            with open(src, "rb") as f:
                h = hashlib.sha256(hdr.digest() + f.read() + " ".join(cflags).encode()).hexdigest()[:24]
            obj = os.path.join(objdir, os.path.basename(src)[:-3] + "." + h + ".o")
            if not os.path.exists(obj):
                cmd = [nvcc] + cflags + ["-c", "-o", obj + ".tmp", src] + (["-Xptxas", "-v"] if verbose else [])
                res = subprocess.run(cmd, capture_output=True, text=True)
                if res.returncode != 0:
                    raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
                os.replace(obj + ".tmp", obj)
                if verbose:
                    print(res.stderr)
            return obj

        with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
            objs = list(ex.map(compile_one, sources()))
        res = subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", SO_PATH] + objs + ["-ldl"],
                             capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("nvcc link failed:\n" + res.stdout + res.stderr)
        if not os.environ.get("GHB_LIB_PATH"):             # the default library: drop the objects of older versions / variants
            keep = set(objs)
            for o in glob.glob(os.path.join(objdir, "*.o")):
                if o not in keep:
                    os.remove(o)
        with open(SO_PATH + ".srchash", "w") as f:
            f.write(_source_hash())
    return SO_PATH


_LIB = None


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    # missing, or edited sources: never run a stale library (an explicit A/B artifact, GHB_LIB_PATH, is used as it is)
    if not os.path.exists(SO_PATH) or (not os.environ.get("GHB_LIB_PATH") and needs_build()):
        build()
    L = ctypes.CDLL(SO_PATH)
    vp, i32, i64, u64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_uint64
    L.ghb_create.argtypes = [i32, ctypes.POINTER(vp)]
    L.ghb_destroy.argtypes = [vp]
    L.ghb_destroy.restype = None
    L.ghb_last_error.argtypes = [vp]
    L.ghb_last_error.restype = ctypes.c_char_p
    L.ghb_set_stream.argtypes = [vp, vp]
    L.ghb_synchronize.argtypes = [vp]
    L.ghb_launch_count.argtypes = [vp]
    L.ghb_launch_count.restype = i64
    L.ghb_plan_kernel_name.argtypes = [vp, i32]
    L.ghb_plan_kernel_name.restype = ctypes.c_char_p
    L.ghb_plan_blocks.argtypes = [vp, i32, vp, vp, i32, vp, i32, vp, ctypes.POINTER(i32)]
    L.ghb_plan_query.argtypes = [vp, i32, ctypes.POINTER(i64)]
    L.ghb_condense_f64.argtypes = [vp, i32, i64, vp, vp, vp, vp, vp, i32]
    L.ghb_restrict_facet_dofs_i64.argtypes = [vp, i64, i32, i32, vp, vp, vp]
    L.ghb_sum_facets_f64.argtypes = [vp, i64, i32, i64, vp, vp]
    L.ghb_l2_projection_dofs_f64.argtypes = [vp, i64, i32, i32, vp, vp, vp, vp]
    L.ghb_expand_records_f64.argtypes = [vp, i32, i64, i32, vp, vp, vp, vp, vp]
    L.ghb_condense_affine_f64.argtypes = [vp, i32, i64, i32, vp, vp, vp, vp, vp, vp, i32]
    L.ghb_condense_assemble_affine_f64.argtypes = [vp, i32, i64, i32, vp, vp, vp, vp, i64, vp, vp, vp]
    L.ghb_backsub_affine_f64.argtypes = [vp, i32, i64, i32, vp, vp, vp, vp, i64, vp, i64, vp, vp, vp]
    L.ghb_assemble_symbolic.argtypes = [vp, i64, i32, vp, i64, ctypes.POINTER(i64)]
    L.ghb_assemble_pattern.argtypes = [vp, vp, vp]
    L.ghb_assemble_numeric_f64.argtypes = [vp, vp, vp, vp, i64, vp, vp]
    L.ghb_assemble_numeric_csr_f64.argtypes = [vp, vp, vp, vp, i64, vp, vp]
    L.ghb_set_option.argtypes = [vp, ctypes.c_char_p, i64]
    L.ghb_condense_scatter_slab_f64.argtypes = [vp, i32, i64, vp, vp, vp, vp, vp, vp, i64, i32]
    L.ghb_assemble_finish_slab_f64.argtypes = [vp, vp, vp, vp, vp, vp, vp]
    L.ghb_condense_scatter_slab_affine_f64.argtypes = [vp, i32, i64, i32, vp, vp, vp, vp, vp, vp, vp, i64, i32]
    L.ghb_device_alloc.argtypes = [vp, i64, ctypes.POINTER(vp)]
    L.ghb_device_free.argtypes = [vp, vp]
    L.ghb_copy.argtypes = [vp, vp, vp, i64]
    L.ghb_host_register.argtypes = [vp, vp, i64]
    L.ghb_host_unregister.argtypes = [vp, vp]
    L.ghb_factors_generation.argtypes = [vp]
    L.ghb_factors_generation.restype = i64
    L.ghb_assemble_current.argtypes = [vp]
    L.ghb_assemble_select.argtypes = [vp, i32]
    L.ghb_assemble_release.argtypes = [vp, i32]
    L.ghb_comm_unique_id.argtypes = [vp, vp]
    L.ghb_comm_init.argtypes = [vp, i32, i32, vp]
    L.ghb_comm_destroy.argtypes = [vp]
    L.ghb_exchange_cut_plane_f64.argtypes = [vp, vp, i64, vp, i64]
    L.ghb_allgather_lambda_f64.argtypes = [vp, vp, vp, vp]
    L.ghb_assemble_symbolic_slab.argtypes = [vp, i64, i64, i32, i32, vp, i64, i64, i64, ctypes.POINTER(i64)]
    L.ghb_pack_cut_plane_f64.argtypes = [vp, i64, i32, i32, vp, vp, vp, vp, vp]
    L.ghb_assemble_numeric_slab_f64.argtypes = [vp, vp, vp, vp, vp, vp, vp]
    L.ghb_condense_assemble_f64.argtypes = [vp, i32, i64, vp, vp, vp, i64, vp, vp, vp]
    L.ghb_backsub_f64.argtypes = [vp, i32, i64, vp, vp, vp, i64, vp, i64, vp, vp, vp]
    L.ghb_scatter_free_dof_values.argtypes = [vp, i32, i64, vp, vp, i64, vp]
    L.ghb_synth_fill_f64.argtypes = [vp, i32, i64, i64, u64, vp, vp]
    L.ghb_cartesian_cell_wise_facets.argtypes = [vp, i32, vp, i64, i64, vp]
    for name in SYMBOLS:
        fn = getattr(L, name)
        if fn.restype is ctypes.c_int and name not in ("ghb_destroy",):
            fn.restype = ctypes.c_int
    _LIB = L
    return L
