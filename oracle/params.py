"""Hyper-parameter transforms (oracle; test infrastructure only).

Follows ``src/util.jl:36-55``: every optimiser parameter is ``exp(p) + 1e-3``.  Callers square the
variances (``optimized.jl:30,136,140``; ``dtc.jl:31,37``; ``temporal_gp_inference.jl:28``) and the
noise (``optimized.jl:34,152``; ``dtc.jl:34-35,44``).
"""
import numpy as np

EQ, MATERN12, MATERN32, MATERN52 = 0, 1, 2, 3
KERNEL_NAMES = {EQ: "EQ", MATERN12: "Matern12", MATERN32: "Matern32", MATERN52: "Matern52"}


def unpack_gp(params):
    """``unpack_gp`` — src/util.jl:36-43.  -> (l, process_var, noise_sigma)."""
    p = np.asarray(params, dtype=np.float64)
    return tuple(float(np.exp(p[i]) + 1e-3) for i in range(3))


def unpack_gpar(params):
    """``unpack_gpar`` — src/util.jl:45-55. -> (time_l, time_var, out_l, out_var, noise_sigma)."""
    p = np.asarray(params, dtype=np.float64)
    return tuple(float(np.exp(p[i]) + 1e-3) for i in range(5))
