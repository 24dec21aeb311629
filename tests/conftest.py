import os
import sys

os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")   # the oracle's cells are small; BLAS threads only fight each other

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
