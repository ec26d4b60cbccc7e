"""gpar-at-scale_b200 — B200-native (sm_100a) GP linear-algebra hot path of GPAR-at-scale.

`csrc/` holds the hand-written CUDA kernels and the C ABI (include/gpar_b200.h) built into
`lib/libgpar_b200.so`; the Python modules here are the host-side mirror of the reference's Julia
interface for this path (same function names, argument meaning and error behaviour), bound to the
ABI with ctypes.  Import it as `gpar_at_scale_b200` (the loader at the repo root maps the hyphenated
directory to that module name).
"""
from ._ffi import (EQ, MATERN12, MATERN32, MATERN52, GparError, PosDefException, load_library, LIB_PATH)
from .context import Context, Group
from . import api, neldermead, parallel, chain

__all__ = ["EQ", "MATERN12", "MATERN32", "MATERN52", "GparError", "PosDefException", "load_library",
           "LIB_PATH", "Context", "Group"]
