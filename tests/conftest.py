import os
import sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
EXAMPLES = os.path.join(ROOT, "examples")        # toy_data.py (mirror of src/data/toy_data.jl) lives with the example drivers
if EXAMPLES not in sys.path:
    sys.path.insert(1, EXAMPLES)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu():
    try:
        import ctypes
        lib = ctypes.CDLL("libcudart.so.12")
        n = ctypes.c_int(0)
        return lib.cudaGetDeviceCount(ctypes.byref(n)) == 0 and n.value > 0
    except OSError:
        try:
            import torch
            return torch.cuda.is_available()
        except Exception:
            return False


HAS_GPU = None


def pytest_collection_modifyitems(config, items):
    global HAS_GPU
    if HAS_GPU is None:
        HAS_GPU = _has_gpu()
    if HAS_GPU:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def ctx():
    """One device context for the GPU tests; creating it fails loudly when the CUDA library is
    missing (no CPU fallback exists)."""
    import gpar_at_scale_b200 as gp
    c = gp.Context(0)
    yield c
    c.close()
