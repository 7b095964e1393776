import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden_addr():
    return np.load(os.path.join(ROOT, "tests", "golden", "polymesh_addr.npz"))


@pytest.fixture(scope="session")
def gpu_ctx():
    """One b200_ctx for the whole GPU session.  Fails loudly (no skip, no fallback) when the CUDA
    library is missing or no device is usable: a GPU test that cannot reach the kernels is a failure."""
    from multiregionfoam_b200 import ldu
    ctx = ldu.Context(device=0)
    yield ctx
    ctx.close()
