import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box)")


def _cuda_device_count():
    """Number of CUDA devices the driver reports (0 without a driver / device); no torch import, no context creation."""
    import ctypes
    try:
        cu = ctypes.CDLL("libcuda.so.1")
        if cu.cuInit(0) != 0:
            return 0
        n = ctypes.c_int(0)
        return n.value if cu.cuDeviceGetCount(ctypes.byref(n)) == 0 else 0
    except OSError:
        return 0


def pytest_collection_modifyitems(config, items):
    """A plain `pytest` on a box without a GPU skips the gpu-marked tests instead of failing in sphb200_create (there is no CPU
    fallback to run them on).  `-m gpu` on such a box still reports them as skipped, never as passed."""
    if _cuda_device_count() > 0:
        return
    skip = pytest.mark.skip(reason="needs a CUDA device (sphb200 has no CPU fallback)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (tests only)."""
    from oracle import oracle as orc
    orc.build()
    return orc


@pytest.fixture(scope="session")
def sphlib():
    """libsphb200.so, built in-tree if necessary (nvcc cross-compiles without a GPU)."""
    from spheral_b200 import build as b
    b.build()
    from spheral_b200 import _lib
    return _lib.lib()
