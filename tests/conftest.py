import os
import sys

# One BLAS thread: the oracle's cells are small (threads only fight each other) and the last bits of LAPACK results
# depend on the thread count, which tests/golden/reference_vectors.npz records.  A pytest plugin may have loaded numpy
# (and OpenBLAS) before this file runs, so the limit is also applied at run time.
os.environ["OPENBLAS_NUM_THREADS"] = "1"
try:
    from threadpoolctl import threadpool_limits
    _BLAS_LIMIT = threadpool_limits(limits=1)
except Exception:       # pragma: no cover
    _BLAS_LIMIT = None

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def small_day():
    from optimalinterpolation_b200.synthetic import make_small_day
    return make_small_day()


@pytest.fixture(scope="session")
def small_oracle(small_day):
    from oracle.gpr_oracle import DayOracle
    return DayOracle.from_day(small_day)
