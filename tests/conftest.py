import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built_libraries():
    """Both shared libraries are built in-tree; build them here if a fresh checkout has none (CPU only: nvcc
    cross-compiles).  On the GPU box the prebuilt files travel with the snapshot."""
    import pecs_b200._lib as L
    if not os.path.exists(L.LIB_PATH):
        from pecs_b200.build import build
        build()
    import oracle
    if not os.path.exists(oracle.LIB_PATH):
        oracle.build()
    yield
