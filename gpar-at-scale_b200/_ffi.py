"""ctypes binding of libgpar_b200.so — the same symbols, argument order and memory layouts the
Julia `ccall` stubs of INTEGRATION.md use (include/gpar_b200.h is the single source of truth).

There is no CPU fallback: if the shared library (or a GPU) is missing, loading / context creation
raises.
"""
import ctypes
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libgpar_b200.so")

GPAR_OK, GPAR_ERR_INVALID, GPAR_ERR_CUDA, GPAR_ERR_NOT_POSDEF, GPAR_ERR_NOMEM = range(5)
EQ, MATERN12, MATERN32, MATERN52 = 0, 1, 2, 3

_c_double_p = ctypes.POINTER(ctypes.c_double)
_c_void_p = ctypes.c_void_p

# symbol -> (restype, argtypes); tests check every symbol of include/gpar_b200.h is exported
SIGNATURES = {
    "gpar_abi_version": (ctypes.c_int, []),
    "gpar_ctx_create": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(_c_void_p)]),
    "gpar_ctx_destroy": (ctypes.c_int, [_c_void_p]),
    "gpar_last_error": (ctypes.c_char_p, [_c_void_p]),
    "gpar_last_timing": (ctypes.c_int, [_c_void_p, _c_double_p, ctypes.POINTER(ctypes.c_int64)]),
    "gpar_last_profile": (ctypes.c_int, [_c_void_p, _c_double_p, ctypes.c_int32]),
    "gpar_measure_peaks": (ctypes.c_int, [_c_void_p, _c_double_p, _c_double_p, _c_double_p]),
    "gpar_dense_bench": (ctypes.c_int, [_c_void_p, ctypes.c_int32, _c_double_p]),
    "gpar_set_inputs": (ctypes.c_int, [_c_void_p, _c_double_p, ctypes.c_int32, ctypes.c_int64]),
    "gpar_set_pseudo": (ctypes.c_int, [_c_void_p, _c_double_p, ctypes.c_int32, ctypes.c_int64]),
    "gpar_set_times": (ctypes.c_int, [_c_void_p, _c_double_p, ctypes.c_int64]),
    "gpar_set_times_range": (ctypes.c_int, [_c_void_p, ctypes.c_double, ctypes.c_double, ctypes.c_int64]),
    "gpar_set_outputs": (ctypes.c_int, [_c_void_p, _c_double_p, ctypes.c_int64, ctypes.c_int32]),
    "gpar_set_noise_vector": (ctypes.c_int, [_c_void_p, _c_double_p, ctypes.c_int64]),
    "gpar_set_merged": (ctypes.c_int, [_c_void_p, _c_double_p, _c_double_p, _c_double_p, ctypes.c_int64, _c_double_p, _c_double_p,
                                       ctypes.c_int64, ctypes.c_int32, ctypes.c_double]),
    "gpar_take_test": (ctypes.c_int, [_c_void_p, _c_double_p, _c_double_p]),
    "gpar_dtc_logpdf": (ctypes.c_int, [_c_void_p, ctypes.c_int, _c_double_p, ctypes.c_int, ctypes.c_double,
                                       _c_double_p, _c_double_p]),
    "gpar_dtc_logpdf_zgrad": (ctypes.c_int, [_c_void_p, ctypes.c_int, _c_double_p, ctypes.c_int, ctypes.c_double,
                                             _c_double_p, _c_double_p, _c_double_p]),
    "gpar_scaled_dtc": (ctypes.c_int, [_c_void_p, ctypes.c_int, ctypes.c_int, _c_double_p, _c_double_p, _c_double_p]),
    "gpar_scaled_dtc_batch": (ctypes.c_int, [_c_void_p, ctypes.c_int, ctypes.c_int, _c_double_p, ctypes.c_int32, _c_double_p, ctypes.POINTER(ctypes.c_int32)]),
    "gpar_scaled_dtc_grad": (ctypes.c_int, [_c_void_p, ctypes.c_int, ctypes.c_int, _c_double_p, _c_double_p, _c_double_p]),
    "gpar_compute_q_u": (ctypes.c_int, [_c_void_p, ctypes.c_int, ctypes.c_int, _c_double_p, _c_double_p,
                                        _c_double_p, _c_double_p]),
    "gpar_sample_q_u": (ctypes.c_int, [_c_void_p, ctypes.c_int, ctypes.c_int, _c_double_p, ctypes.c_uint64, ctypes.c_int32,
                                       _c_double_p, _c_double_p]),
    "gpar_scaled_predict": (ctypes.c_int, [_c_void_p, ctypes.c_int, ctypes.c_int, _c_double_p, _c_double_p, ctypes.c_int32,
                                           _c_double_p, _c_double_p]),
    "gpar_lgssm_logpdf": (ctypes.c_int, [_c_void_p, ctypes.c_int, _c_double_p, ctypes.c_int32, _c_double_p]),
    "gpar_lgssm_logpdf_grad": (ctypes.c_int, [_c_void_p, ctypes.c_int, _c_double_p, ctypes.c_int32, _c_double_p, _c_double_p]),
    "gpar_lgssm_decorrelate": (ctypes.c_int, [_c_void_p, ctypes.c_int, _c_double_p, _c_double_p, _c_double_p]),
    "gpar_lgssm_smooth": (ctypes.c_int, [_c_void_p, ctypes.c_int, _c_double_p, _c_double_p, _c_double_p, _c_double_p]),
    "gpar_exact_logpdf": (ctypes.c_int, [_c_void_p, ctypes.c_int, ctypes.c_int, _c_double_p, ctypes.c_int32, _c_double_p]),
    "gpar_exact_logpdf_batch": (ctypes.c_int, [_c_void_p, ctypes.c_int, ctypes.c_int, _c_double_p, ctypes.c_int32, ctypes.c_int32, _c_double_p,
                                               ctypes.POINTER(ctypes.c_int32)]),
    "gpar_exact_posterior": (ctypes.c_int, [_c_void_p, ctypes.c_int, ctypes.c_int, _c_double_p, ctypes.c_int32,
                                            _c_double_p, ctypes.c_int64, _c_double_p, _c_double_p]),
    "gpar_group_create": (ctypes.c_int, [ctypes.POINTER(ctypes.c_int32), ctypes.c_int32, ctypes.POINTER(_c_void_p)]),
    "gpar_group_destroy": (ctypes.c_int, [_c_void_p]),
    "gpar_group_size": (ctypes.c_int32, [_c_void_p]),
    "gpar_group_ctx": (_c_void_p, [_c_void_p, ctypes.c_int32]),
    "gpar_group_last_error": (ctypes.c_char_p, [_c_void_p]),
    "gpar_group_dtc_logpdf": (ctypes.c_int, [_c_void_p, ctypes.c_int, _c_double_p, ctypes.c_int, ctypes.c_double,
                                             _c_double_p, _c_double_p, ctypes.POINTER(ctypes.c_int32)]),
    "gpar_group_scaled_dtc": (ctypes.c_int, [_c_void_p, ctypes.c_int, ctypes.c_int, _c_double_p,
                                             _c_double_p, _c_double_p, ctypes.POINTER(ctypes.c_int32)]),
    "gpar_group_dtc_logpdf_sharded": (ctypes.c_int, [_c_void_p, ctypes.c_int, _c_double_p, ctypes.c_int, ctypes.c_double, _c_double_p, _c_double_p]),
    "gpar_group_scaled_dtc_sharded": (ctypes.c_int, [_c_void_p, ctypes.c_int, ctypes.c_int, _c_double_p, ctypes.POINTER(ctypes.c_int64), _c_double_p, _c_double_p]),
    "gpar_scaled_slice_begin": (ctypes.c_int, [_c_void_p, ctypes.c_int, ctypes.c_int, _c_double_p, ctypes.c_int64, ctypes.c_int32,
                                               ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int64)]),
    "gpar_scaled_slice_summary": (ctypes.c_int, [_c_void_p, _c_void_p]),
    "gpar_scaled_slice_stats": (ctypes.c_int, [_c_void_p, _c_void_p, ctypes.c_int32, _c_void_p]),
    "gpar_scaled_slice_value": (ctypes.c_int, [_c_void_p, _c_void_p, _c_double_p]),
    "gpar_scaled_slice_tangent_summary": (ctypes.c_int, [_c_void_p, _c_void_p, _c_void_p]),
    "gpar_scaled_slice_grad_partial": (ctypes.c_int, [_c_void_p, _c_void_p, ctypes.c_int32, _c_double_p]),
    "gpar_scaled_slice_grad_finish": (ctypes.c_int, [_c_void_p, _c_double_p, _c_double_p, _c_double_p]),
    "gpar_group_fit": (ctypes.c_int, [_c_void_p, _c_double_p, ctypes.c_int64, _c_void_p, ctypes.c_int32, ctypes.c_int, ctypes.c_int,
                                      ctypes.c_int32, ctypes.c_int32, _c_double_p, _c_double_p, ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int32)]),
    "gpar_group_compute_q_u_sharded": (ctypes.c_int, [_c_void_p, ctypes.c_int, ctypes.c_int, _c_double_p, ctypes.POINTER(ctypes.c_int64), _c_double_p, _c_double_p, _c_double_p]),
    "gpar_group_sample_q_u_sharded": (ctypes.c_int, [_c_void_p, ctypes.c_int, ctypes.c_int, _c_double_p, ctypes.POINTER(ctypes.c_int64), ctypes.c_uint64, ctypes.c_int32,
                                                     _c_double_p, _c_double_p]),
    "gpar_group_fit_sharded": (ctypes.c_int, [_c_void_p, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int64), _c_double_p, ctypes.c_int32, ctypes.c_int32,
                                              _c_double_p, _c_double_p, ctypes.POINTER(ctypes.c_int32)]),
    "gpar_group_broadcast": (ctypes.c_int, [_c_void_p, ctypes.c_int32, _c_double_p, ctypes.c_int64, _c_double_p]),
    "gpar_set_inputs_column": (ctypes.c_int, [_c_void_p, ctypes.c_int32, _c_double_p]),
    "gpar_set_merged_test_column": (ctypes.c_int, [_c_void_p, ctypes.c_int32, _c_double_p]),
}


class FitTask(ctypes.Structure):
    """struct gpar_fit_task (include/gpar_b200.h)"""
    _fields_ = [("X", _c_double_p), ("D", ctypes.c_int32), ("Z", _c_double_p), ("M", ctypes.c_int64), ("y", _c_double_p),
                ("theta0", ctypes.c_double * 5)]

_lib = None


class GparError(RuntimeError):
    def __init__(self, status, message):
        super().__init__("libgpar_b200 status %d: %s" % (status, message))
        self.status = status


class PosDefException(GparError):
    """Raised for GPAR_ERR_NOT_POSDEF — the Julia shim throws LinearAlgebra.PosDefException, which
    is what `cholesky` does in the reference (src/gp/dtc.jl:119-120)."""


def load_library(path=None):
    """dlopen the in-tree CUDA library.  Raises OSError when it has not been built — the product
    path must fail loudly rather than fall back to anything on the CPU."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or os.environ.get("GPAR_B200_LIB") or LIB_PATH     # GPAR_B200_LIB: tuning builds of the same library
    if not os.path.exists(p):
        raise OSError("libgpar_b200.so not found at %s — build it with `make` or __graft_entry__.build()" % p)
    lib = ctypes.CDLL(p, mode=ctypes.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _lib = lib
    return lib


def as_f64(a, order="C"):
    return np.require(a, dtype=np.float64, requirements=["C" if order == "C" else "F", "A"])


def dptr(a):
    return a.ctypes.data_as(_c_double_p) if a is not None else None
