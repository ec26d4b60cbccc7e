"""ctypes binding of oracle/c/oracle_kalman.c (TEST INFRASTRUCTURE / CPU BASELINE ONLY).
Same semantics as oracle/lgssm.py, usable at large N.  `build()` compiles it with gcc."""
import ctypes
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle_c.so")
_lib = None
_dp = ctypes.POINTER(ctypes.c_double)


def build(force=False):
    src = os.path.join(_HERE, "c", "oracle_kalman.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        subprocess.check_call(["gcc", "-O3", "-march=x86-64-v3", "-fPIC", "-shared", "-fopenmp", "-o", _SO, src, "-lm"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.oracle_kalman_filter.restype = ctypes.c_double
        _lib.oracle_kalman_filter.argtypes = [ctypes.c_int, ctypes.c_long, _dp, _dp, _dp, ctypes.c_double, ctypes.c_double,
                                               ctypes.c_double, _dp, _dp, _dp, _dp, _dp]
        _lib.oracle_kalman_smooth.restype = ctypes.c_double
        _lib.oracle_kalman_smooth.argtypes = [ctypes.c_int, ctypes.c_long, _dp, _dp, _dp, ctypes.c_double, ctypes.c_double,
                                               ctypes.c_double, _dp, _dp]
        _lib.oracle_kalman_filter_batch.restype = None
        _lib.oracle_kalman_filter_batch.argtypes = [ctypes.c_int, ctypes.c_long, ctypes.c_long, _dp, _dp, _dp, _dp, _dp, _dp,
                                                     ctypes.c_long, _dp, _dp]
        _lib.oracle_kalman_smooth_batch.restype = None
        _lib.oracle_kalman_smooth_batch.argtypes = [ctypes.c_int, ctypes.c_long, ctypes.c_long, _dp, _dp, _dp, ctypes.c_double,
                                                     ctypes.c_double, ctypes.c_double, _dp, _dp, _dp]
        _lib.oracle_decorrelate_columns.restype = None
        _lib.oracle_decorrelate_columns.argtypes = [ctypes.c_int, ctypes.c_long, ctypes.c_long, _dp, _dp, _dp, ctypes.c_double,
                                                     ctypes.c_double, ctypes.c_double, _dp]
    return _lib


def _p(a):
    return a.ctypes.data_as(_dp) if a is not None else None


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _noise(noise):
    if np.ndim(noise) == 0:
        return None, float(noise)
    return _f(noise), 0.0


def kalman_decorrelate(kind, t, y, l, s, noise):
    """C twin of oracle.lgssm.kalman_decorrelate -> (lml, alpha)."""
    t = _f(t); y = _f(y); rv, nz = _noise(noise)
    alpha = np.zeros(len(y))
    lml = lib().oracle_kalman_filter(kind, len(y), _p(t), _p(y), _p(rv), nz, float(l), float(s), _p(alpha), None, None, None, None)
    return lml, alpha


def kalman_logpdf(kind, t, y, l, s, noise):
    t = _f(t); y = _f(y); rv, nz = _noise(noise)
    return lib().oracle_kalman_filter(kind, len(y), _p(t), _p(y), _p(rv), nz, float(l), float(s), None, None, None, None, None)


def kalman_smooth(kind, t, y, l, s, noise):
    """C twin of oracle.lgssm.kalman_smooth -> (lml, mean, var)."""
    t = _f(t); y = _f(y); rv, nz = _noise(noise)
    mean = np.zeros(len(y)); var = np.zeros(len(y))
    lml = lib().oracle_kalman_smooth(kind, len(y), _p(t), _p(y), _p(rv), nz, float(l), float(s), _p(mean), _p(var))
    return lml, mean, var


def kalman_filter_batch(kind, t, Y, l, s, noise, rvec=None, want_alpha=False):
    """Y: (batch, n); l, s, noise scalars (shared model) or (batch,) arrays (independent models)."""
    t = _f(t); Y = _f(np.atleast_2d(Y)); b, n = Y.shape
    l = _f(np.atleast_1d(l)); s = _f(np.atleast_1d(s)); nz = _f(np.atleast_1d(noise))
    rv = _f(rvec) if rvec is not None else None
    alpha = np.zeros((b, n)) if want_alpha else None
    lml = np.zeros(b)
    lib().oracle_kalman_filter_batch(kind, n, b, _p(t), _p(Y), _p(rv), _p(nz), _p(l), _p(s), len(l), _p(alpha), _p(lml))
    return (lml, alpha) if want_alpha else lml


def kalman_smooth_batch(kind, t, Y, l, s, noise):
    t = _f(t); Y = _f(np.atleast_2d(Y)); b, n = Y.shape
    rv, nz = _noise(noise)
    mean = np.zeros((b, n)); var = np.zeros((b, n)); lml = np.zeros(b)
    lib().oracle_kalman_smooth_batch(kind, n, b, _p(t), _p(Y), _p(rv), nz, float(l), float(s), _p(mean), _p(var), _p(lml))
    return lml, mean, var


def decorrelate_columns(kind, t, C, l, s, noise):
    """The reference's beta loop (dtc.jl:108-117) over the columns of C (n x m)."""
    t = _f(t); Cf = np.asfortranarray(C, dtype=np.float64); n, m = Cf.shape
    rv, nz = _noise(noise)
    beta = np.zeros((n, m), order="F")
    lib().oracle_decorrelate_columns(kind, n, m, _p(t), Cf.ctypes.data_as(_dp), _p(rv), nz, float(l), float(s),
                                     beta.ctypes.data_as(_dp))
    return beta
