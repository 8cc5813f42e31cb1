"""ctypes binding of libfinegpu.so (include/fegpu.h).  This is the Python twin of the `ccall` layer of
julia/FinEtoolsGPU.jl.  There is NO fallback: if the shared library is missing or no GPU is present, calls raise."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfinegpu.so")

c_i64p = C.POINTER(C.c_int64)
c_f64p = C.POINTER(C.c_double)
c_i32p = C.POINTER(C.c_int32)
VP = C.c_void_p

# every symbol include/fegpu.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "fegpu_create": (C.c_int32, [C.POINTER(VP), C.c_int32]),
    "fegpu_destroy": (C.c_int32, [VP]),
    "fegpu_last_error": (C.c_char_p, [VP]),
    "fegpu_set_stream": (C.c_int32, [VP, VP]),
    "fegpu_set_async": (C.c_int32, [VP, C.c_int32]),
    "fegpu_set_overlap": (C.c_int32, [VP, C.c_int32]),
    "fegpu_synchronize": (C.c_int32, [VP]),
    "fegpu_cache_release": (C.c_int32, [VP]),
    "fegpu_launch_count": (C.c_int64, [VP]),
    "fegpu_measure_peaks": (C.c_int32, [VP, c_f64p, c_f64p]),
    "fegpu_host_alloc": (C.c_int32, [C.POINTER(VP), C.c_int64]),
    "fegpu_host_free": (C.c_int32, [VP]),
    "fegpu_marks_begin": (C.c_int32, [VP]),
    "fegpu_marks_read": (C.c_int32, [VP, VP, C.c_int64]),
    "fegpu_geom_update_window": (C.c_int32, [VP, VP]),
    "fegpu_mesh_window": (C.c_int32, [VP, c_i64p, c_i64p, c_i64p]),
    "fegpu_mesh_upload": (C.c_int32, [VP, C.c_int32, C.c_int64, VP, C.c_int64, C.c_int32, VP, C.POINTER(VP)]),
    "fegpu_mesh_destroy": (C.c_int32, [VP]),
    "fegpu_geom_update": (C.c_int32, [VP, VP]),
    "fegpu_rule_set": (C.c_int32, [VP, C.c_int32, VP, VP, VP]),
    "fegpu_otherdimension_set": (C.c_int32, [VP, C.c_double]),
    "fegpu_csys_set": (C.c_int32, [VP, VP]),
    "fegpu_partition_set": (C.c_int32, [VP, VP, C.c_int32]),
    "fegpu_dofmap_upload": (C.c_int32, [VP, VP, C.c_int32, VP, C.c_int64, C.c_int64, C.POINTER(VP)]),
    "fegpu_dofmap_destroy": (C.c_int32, [VP]),
    "fegpu_asm_create": (C.c_int32, [VP, C.POINTER(VP)]),
    "fegpu_asm_destroy": (C.c_int32, [VP]),
    "fegpu_asm_set_symmetric": (C.c_int32, [VP, C.c_int32]),
    "fegpu_asm_set_lumping": (C.c_int32, [VP, C.c_int32]),
    "fegpu_bilform_diffusion": (C.c_int32, [VP, VP, C.c_int32, VP, VP]),
    "fegpu_bilform_lin_elastic": (C.c_int32, [VP, VP, VP, VP]),
    "fegpu_bilform_dot": (C.c_int32, [VP, VP, VP, C.c_int32, C.c_double, VP]),
    "fegpu_bilform_convection": (C.c_int32, [VP, VP, VP, C.c_double, VP]),
    "fegpu_bilform_div_grad": (C.c_int32, [VP, VP, C.c_double, VP]),
    "fegpu_bilform_masslike": (C.c_int32, [VP, VP, VP, C.c_int32, C.c_double, VP]),
    "fegpu_linform_dot": (C.c_int32, [VP, VP, VP, C.c_int32, C.c_double, VP]),
    "fegpu_vec_startassembly": (C.c_int32, [VP, C.c_int64]),
    "fegpu_vec_assemble": (C.c_int32, [VP, VP, VP, C.c_int64]),
    "fegpu_makevector": (C.c_int32, [VP]),
    "fegpu_makevector_size": (C.c_int32, [VP, c_i64p]),
    "fegpu_makevector_copy": (C.c_int32, [VP, VP]),
    "fegpu_startassembly": (C.c_int32, [VP, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int64]),
    "fegpu_assemble": (C.c_int32, [VP, VP, VP, C.c_int64, VP, C.c_int64]),
    "fegpu_triplets_append": (C.c_int32, [VP, C.c_int64, VP, VP, VP]),
    "fegpu_makematrix": (C.c_int32, [VP]),
    "fegpu_makematrix_sizes": (C.c_int32, [VP, c_i64p, c_i64p, c_i64p]),
    "fegpu_makematrix_copy": (C.c_int32, [VP, VP, VP, VP]),
    "fegpu_makematrix_view": (C.c_int32, [VP, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int32]),
    "fegpu_transfer_stats": (C.c_int32, [VP, c_i64p, c_i64p]),
    "fegpu_transfer_compressed": (C.c_int32, [VP, c_i64p]),
    "fegpu_transfer_stenciled": (C.c_int32, [VP, c_i64p]),
    "fegpu_makematrix_copy_values": (C.c_int32, [VP, VP]),
    "fegpu_makematrix_device": (C.c_int32, [VP, C.POINTER(VP), C.POINTER(VP), C.POINTER(VP)]),
    "fegpu_coo_copy": (C.c_int32, [VP, VP, VP, VP, VP, VP]),
    "fegpu_block_counts": (C.c_int32, [VP, VP]),
    "fegpu_gather_plan": (C.c_int32, [VP, VP, C.c_int32, C.c_int64, VP, VP]),
    "fegpu_gather_place": (C.c_int32, [VP, VP, C.c_int32, C.c_int32, C.c_int64, VP, VP, VP, VP, VP]),
    "fegpu_gather_sort_columns": (C.c_int32, [VP, C.c_int64, VP, VP, VP, c_i64p]),
    "fegpu_last_timings": (C.c_int32, [VP, c_f64p]),
    "fegpu_pattern_was_cached": (C.c_int32, [VP]),
    "fegpu_pattern_invalidate": (C.c_int32, [VP]),
    "fegpu_pattern_path": (C.c_int32, [VP]),
}

_lib = None


class FEGPUError(RuntimeError):
    """Mirrors the reference's `error("...")` (ErrorException); `.code` is the C status."""

    def __init__(self, code, msg):
        super().__init__(msg)
        self.code = code


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FEGPUError(-1, "libfinegpu.so is not built (%s); run `python -c 'import __graft_entry__ as g; g.build()'` -- "
                             "there is no CPU fallback" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            f = getattr(L, name)
            f.restype = res
            f.argtypes = args
        _lib = L
    return _lib


def check(status, ctx=None):
    if status != 0:
        msg = lib().fegpu_last_error(ctx)
        raise FEGPUError(status, (msg or b"").decode() or ("fegpu error %d" % status))


def fptr(a):
    return a.ctypes.data_as(VP) if a is not None else None


def colmajor_f64(a):
    """Bytes of a (Julia) column-major Float64 matrix from a NumPy array."""
    return np.asfortranarray(np.asarray(a, dtype=np.float64))


def colmajor_i64(a):
    return np.asfortranarray(np.asarray(a, dtype=np.int64))
