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
    # always: both builds are incremental by mtime, so a stale library is never tested against new sources (on the GPU
    # box nvcc exists too; if a toolchain is missing there the prebuilt files that travelled with the snapshot are used)
    import pecs_b200._lib as L
    try:
        from pecs_b200.build import build
        build()
    except Exception:
        if not os.path.exists(L.LIB_PATH):
            raise
    import oracle
    try:
        oracle.build()
    except Exception:
        if not os.path.exists(oracle.LIB_PATH):
            raise
    yield
