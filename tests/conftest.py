import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def orc():
    from oracle import pyoracle as po
    return po.oracle()

# numpy's BLAS thread pool and torch's OpenMP runtime share the test process; a multi-threaded
# BLAS call after torch.distributed / fork-based tests has been seen to deadlock.  The linear
# algebra in the tests is tiny: run it single-threaded.
try:
    from threadpoolctl import threadpool_limits
    _BLAS_LIMIT = threadpool_limits(limits=1, user_api="blas")
except Exception:  # pragma: no cover
    _BLAS_LIMIT = None
