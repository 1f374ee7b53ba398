import os
import sys

import pytest

os.environ.setdefault("DDS_COMM_TIMEOUT_S", "40")  # a wedged rank should fail a test in seconds, not minutes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _ensure_built():
    """the native pieces are built in-tree by __graft_entry__.build(); if this checkout has not been built yet
    (fresh clone: *.so is git-ignored), build now rather than fail every test at import"""
    lib = os.path.join(ROOT, "ddstore_b200", "libddstore_b200.so")
    cy = [f for f in os.listdir(os.path.join(ROOT, "ddstore_b200", "cython")) if f.startswith("pyddstore") and f.endswith(".so")]
    if not os.path.exists(lib) or not cy or not os.path.exists(os.path.join(ROOT, "oracle", "liboracle.so")):
        import __graft_entry__
        __graft_entry__.build()


def pytest_configure(config):
    _ensure_built()
    config.addinivalue_line("markers", "gpu: needs a real B200 (run under gpurun / by the driver)")
    config.addinivalue_line("markers", "multigpu: needs >= 2 GPUs on the box")


@pytest.fixture(scope="session")
def coracle():
    from oracle.oracle import COracle
    return COracle()
