import os
import sys

import pytest

os.environ.setdefault("DDS_COMM_TIMEOUT_S", "40")  # a wedged rank should fail a test in seconds, not minutes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run under gpurun / by the driver)")
    config.addinivalue_line("markers", "multigpu: needs >= 2 GPUs on the box")


@pytest.fixture(scope="session")
def coracle():
    from oracle.oracle import COracle
    return COracle()
