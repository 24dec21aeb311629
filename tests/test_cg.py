"""CPU: the resumable optimiser state machine (csrc/cg_scipy.h, host build) against
scipy.optimize.minimize(method='CG', jac=True) -- the call the reference makes at GPR_CS2S3.py:166.
The sequence of evaluation points must be bit-identical."""
import warnings

import numpy as np
import pytest
import scipy.optimize

from cg_driver import minimize_cg


def _compare(fun, x0):
    xs = []

    def rec(x):
        xs.append(np.array(x))
        return fun(x)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        r = scipy.optimize.minimize(rec, list(x0), jac=True, method="CG")
        m = minimize_cg(fun, x0)
    assert len(xs) == len(m["trace"]) == r.nfev == m["nfev"]
    for a, b in zip(xs, m["trace"]):
        assert np.array_equal(a, b, equal_nan=True)
    assert np.array_equal(r.x, m["x"], equal_nan=True)
    assert r.status == m["status"] and r.nit == m["nit"]
    assert (r.fun == m["fun"]) or (np.isnan(r.fun) and np.isnan(m["fun"]))
    return r


def rosen(x):
    return scipy.optimize.rosen(x), scipy.optimize.rosen_der(x)


@pytest.mark.parametrize("x0", [[-1.2, 1, 0.5, 2, 1], [3, -2, 1, 0, 4, 1], [0.1, 0.2, 0.3, 0.4, 0.5]])
def test_rosenbrock_trace_identical(x0):
    r = _compare(rosen, x0)
    assert r.status == 0


def test_quadratic_and_nonfinite():
    A = np.diag([1.0, 10.0, 100.0, 0.1, 5.0])
    _compare(lambda x: (0.5 * x @ A @ x, A @ x), [1.0, -2.0, 3.0, 0.5, -1.0])
    # objective that turns inf far from the origin (SMLII returns inf on Cholesky failure)
    def f(x):
        if np.abs(x).max() > 3.0:
            return np.inf, np.ones(5) * np.inf
        return float(np.sum(np.cosh(x)) - 0.3 * x[0]), np.sinh(x) - np.r_[0.3, 0, 0, 0, 0]
    _compare(f, [2.5, -2.5, 1.0, 0.0, 2.9])


@pytest.mark.parametrize("cell", [5, 100, 222, 300, 512, 600])
def test_smlii_trace_identical(small_day, small_oracle, cell):
    """On the reference objective itself (incl. its 2x gradient components): runs ending with
    scipy status 0, 2 (precision loss) and 3 (NaN) are all reproduced evaluation by evaluation."""
    from oracle.gpr_oracle import nlml_grad
    _, inp, out, _ = small_oracle.cell_data(cell)
    mX = np.ones(len(out)) * small_day.mean
    _compare(lambda x: nlml_grad(x, inp, out, mX), small_day.x0)
