import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box)")


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
