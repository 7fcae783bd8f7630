import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
SEED = 20261017


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


@pytest.fixture(scope="session")
def lib_built():
    """The CUDA library, built in-tree (nvcc cross-compiles without a GPU)."""
    from fidibench_b200 import build
    return build.build()


@pytest.fixture(scope="session")
def fb(lib_built):
    import fidibench_b200
    return fidibench_b200


@pytest.fixture(scope="session")
def gpu_fb(fb):
    """fidibench_b200 with a usable device.  GPU tests must FAIL, not skip, if the
    CUDA path cannot run: a silent skip would read as a fallback."""
    n = fb.device_count()
    assert n >= 1, "no CUDA device visible to libfidib200.so"
    return fb
