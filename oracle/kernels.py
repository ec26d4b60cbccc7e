"""Covariance kernels and input packing (oracle; test infrastructure only).

Stheno.jl 0.6 semantics [un-vendored dependency]: ``kernel(k; l, s) = s * stretch(k, 1/l)``
(cf. the explicit form ``process_var^2 * stretch(kernel_structure, 1/l)`` at
``src/gp/optimized.jl:30-31``); base kernels with r = ||x - x'||_2:
EQ exp(-r^2/2); Matern12 exp(-r); Matern32 (1+sqrt3 r)exp(-sqrt3 r);
Matern52 (1+sqrt5 r+5r^2/3)exp(-sqrt5 r).
"""
import numpy as np
from .params import EQ, MATERN12, MATERN32, MATERN52


def to_colvecs(inputs):
    """``to_ColVecs`` — src/util.jl:16-31.  A list of per-feature 1-D vectors becomes the D x N
    column-major matrix of a Stheno ``ColVecs``; we return its memory image, an (N, D) C-contiguous
    array (N records of D contiguous doubles), which is what the C ABI receives."""
    if isinstance(inputs, np.ndarray) and inputs.ndim == 2:
        return np.ascontiguousarray(inputs, dtype=np.float64)
    cols = [np.asarray(c, dtype=np.float64).ravel() for c in inputs]
    return np.ascontiguousarray(np.stack(cols, axis=1))


def base_kernel(kind, r):
    """Base kernel as a function of the (already length-scaled) distance r >= 0."""
    r = np.asarray(r, dtype=np.float64)
    if kind == EQ:
        return np.exp(-0.5 * r * r)
    if kind == MATERN12:
        return np.exp(-r)
    if kind == MATERN32:
        a = np.sqrt(3.0) * r
        return (1.0 + a) * np.exp(-a)
    if kind == MATERN52:
        a = np.sqrt(5.0) * r
        return (1.0 + a + a * a / 3.0) * np.exp(-a)
    raise ValueError("unknown kernel %r" % (kind,))


def _dist(X, Z):
    X = np.atleast_2d(np.asarray(X, dtype=np.float64))
    Z = np.atleast_2d(np.asarray(Z, dtype=np.float64))
    d2 = np.zeros((X.shape[0], Z.shape[0]))
    for j in range(X.shape[1]):  # direct (x - z)^2 sum: no cancellation
        diff = X[:, j][:, None] - Z[:, j][None, :]
        d2 += diff * diff
    return np.sqrt(d2)


def pairwise(kind, X, Z, l=1.0, s=1.0):
    """``pairwise(kernel(k; l, s), X, Z)`` — call sites src/gp/dtc.jl:104 (``cov(f,u)``),
    src/gp/gpar_scaled_inference.jl:89,156,157.  X: (N, D), Z: (M, D) records.  No noise."""
    X = np.asarray(X, dtype=np.float64)
    Z = np.asarray(Z, dtype=np.float64)
    if X.ndim == 1:
        X = X[:, None]
    if Z.ndim == 1:
        Z = Z[:, None]
    return s * base_kernel(kind, _dist(X, Z) / l)


scaled_kernel_matrix = pairwise


def get_time_mask(input_length):
    """src/util.jl:102-106."""
    m = np.zeros(input_length)
    m[0] = 1.0
    return m


def get_output_mask(input_length):
    """src/util.jl:111-123 (throws DomainError for input_length <= 1)."""
    if input_length <= 1:
        raise ValueError("Input length must be integer greater than 1")
    m = np.zeros((input_length - 1, input_length))
    for row in range(input_length - 1):
        m[row, row + 1] = 1.0
    return m


def stretched_pairwise(kind, X, Z, mask):
    """``pairwise(stretch(k, mask), X, Z)``: vector mask -> r = |mask.(x-z)|; matrix mask ->
    r = ||mask (x-z)||  (src/util.jl:57-96 docstring example)."""
    X = np.asarray(X, dtype=np.float64)
    Z = np.asarray(Z, dtype=np.float64)
    mask = np.asarray(mask, dtype=np.float64)
    if mask.ndim == 1:
        return base_kernel(kind, np.abs((X @ mask)[:, None] - (Z @ mask)[None, :]))
    return base_kernel(kind, _dist(X @ mask.T, Z @ mask.T))


def gpar_kernel_matrix(k_time, k_out, X, Z, time_l, time_var, out_l, out_var):
    """Exact-GPAR kernel — ``create_gpar_kernel`` src/gp/optimized.jl:132-144:
    time_var^2 k_t(|dx_1|/time_l) + out_var^2 k_o(||dx_{2:D}||/out_l)."""
    X = np.asarray(X, dtype=np.float64)
    Z = np.asarray(Z, dtype=np.float64)
    kt = pairwise(k_time, X[:, :1], Z[:, :1], l=time_l, s=time_var ** 2)
    ko = pairwise(k_out, X[:, 1:], Z[:, 1:], l=out_l, s=out_var ** 2)
    return kt + ko
