import os
import sys

import pytest

# the tests run the warp-specialised kernels with their mbarrier watchdog compiled in: a protocol error traps with the
# name of the stuck barrier instead of hanging the device (the lean variants are what bench.py times)
os.environ.setdefault("DESIRE_GRU3_WATCHDOG", "1")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def lib():
    from desire_b200 import _lib
    return _lib.load()
