"""CPU: the fast-mode optimiser (csrc/lbfgs_fast.h, host build) -- limited-memory BFGS with the More'-Thuente line search and
scipy L-BFGS-B's stopping rules.  It is NOT the parity mode (the reference calls scipy CG, tests/test_cg.py); what is checked
is what BASELINE.json's north_star (5) needs from it: it reaches scipy L-BFGS-B's optimum (or a lower one) in a comparable
number of evaluations, survives inf-returning objectives, and on the reference's own objective (true gradient) ends at or
below the NLML the reference's CG reaches, with several times fewer evaluations."""
import warnings

import numpy as np
import pytest
import scipy.optimize

from cg_driver import minimize_lbfgs


def rosen(x):
    return scipy.optimize.rosen(x), scipy.optimize.rosen_der(x)


@pytest.mark.parametrize("x0", [[-1.2, 1, 0.5, 2, 1], [3, -2, 1, 0, 4, 1], [0.1, 0.2, 0.3, 0.4, 0.5]])
def test_rosenbrock_matches_lbfgsb(x0):
    m = minimize_lbfgs(rosen, x0)
    r = scipy.optimize.minimize(rosen, x0, jac=True, method="L-BFGS-B")
    assert m["status"] == 0
    assert m["fun"] <= r.fun + 1e-8
    assert m["nfev"] <= 1.5 * r.nfev + 5


def test_quadratic_and_nonfinite():
    A = np.diag([1.0, 10.0, 100.0, 0.1, 5.0])
    m = minimize_lbfgs(lambda x: (0.5 * x @ A @ x, A @ x), [1.0, -2.0, 3.0, 0.5, -1.0])
    assert m["status"] == 0 and m["fun"] < 1e-9 and m["nfev"] < 40

    def f(x):      # turns inf far from the origin, like SMLII on a Cholesky failure (GPR_CS2S3.py:139-140)
        if np.abs(x).max() > 3.0:
            return np.inf, np.ones(5) * np.inf
        return float(np.sum(np.cosh(x)) - 0.3 * x[0]), np.sinh(x) - np.r_[0.3, 0, 0, 0, 0]
    m = minimize_lbfgs(f, [2.5, -2.5, 1.0, 0.0, 2.9])
    assert m["status"] == 0 and abs(m["fun"] - (5 - 0.3 * np.arcsinh(0.3) + np.cosh(np.arcsinh(0.3)) - 1)) < 1e-8
    m = minimize_lbfgs(lambda x: (np.nan, np.full(5, np.nan)), [0.0] * 5)
    assert m["status"] == 3 and m["nfev"] == 1


@pytest.mark.parametrize("cell", [5, 100, 222, 512])
def test_smlii_reaches_reference_or_lower(small_day, small_oracle, cell):
    from oracle.gpr_oracle import nlml_grad
    _, inp, out, _ = small_oracle.cell_data(cell, sort=True)
    mX = np.ones(len(out)) * small_day.mean

    def exact(x):      # the reference's components 3 and 4 are twice the true derivative
        f, g = nlml_grad(x, inp, out, mX)
        g = g.copy(); g[3] /= 2; g[4] /= 2
        return f, g
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = minimize_lbfgs(exact, small_day.x0)
        r = scipy.optimize.minimize(exact, small_day.x0, jac=True, method="L-BFGS-B")
        c = scipy.optimize.minimize(lambda x: nlml_grad(x, inp, out, mX), small_day.x0, jac=True, method="CG")
    assert m["status"] in (0, 2)
    assert m["fun"] <= r.fun + 1e-6 * abs(r.fun)
    if np.isfinite(c.fun):
        assert m["fun"] <= float(c.fun) + 1e-6 * abs(float(c.fun))
    assert m["nfev"] < c.nfev
