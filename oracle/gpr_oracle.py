"""CPU restatement (numpy/scipy) of the reference's per-cell GP hot path.

TEST INFRASTRUCTURE ONLY — see oracle/__init__.py.  Every function cites the lines of
/root/reference/2021_paper_production/GPR_CS2S3.py it restates.  The same numpy/scipy
primitives are used in the same order as the reference (pdist/squareform/cdist,
np.linalg.cholesky, np.linalg.solve, scipy.optimize.minimize(method='CG')), so results are
bit-identical to the reference functions on the same library versions; that identity is
what tests/test_oracle.py checks in the build container and what tests/golden/ pins for the
GPU box (where /root/reference does not exist).

Third-party arithmetic the reference relies on (unpinned there; versions used here are
recorded in tests/golden/*.npz): numpy.linalg (LAPACK dpotrf/dgesv), scipy.spatial
(cKDTree, pdist/cdist), scipy.optimize (Polak-Ribiere+ CG with DCSRCH line search).
"""
from __future__ import annotations

import numpy as np
import scipy.optimize
import scipy.spatial
from scipy.spatial.distance import cdist, pdist, squareform

_ROOT3 = np.sqrt(3.)
_LOG2PI = np.log(2 * np.pi)


def matern32(x, ell, sf2, xs=None, want_grad=False):
    """Matern-3/2 ARD covariance, GPR_CS2S3.py:78-105.

    k = sf2*(1+Q)exp(-Q), Q = || sqrt(3) x_i/ell - sqrt(3) x_j/ell ||_2      (:93-94)
    dk[theta] = sf2 * q_theta^2 exp(-Q), q_theta = sqrt(3)|x_i,theta - x_j,theta|/ell_theta (:96-98)
    With ``xs`` the cross-covariance (n x ns) is returned instead (:100-101).
    """
    ell = list(ell)
    if xs is not None:
        Q = cdist(_ROOT3 * x / ell, _ROOT3 * xs / ell, 'euclidean')
        return sf2 * ((1 + Q) * np.exp(-Q))
    Q = squareform(pdist(_ROOT3 * x / ell, 'euclidean'))
    eQ = np.exp(-Q)
    k = (1 + Q) * eQ
    if not want_grad:
        return sf2 * k
    dk = np.zeros((len(ell),) + k.shape)
    for th in range(len(ell)):
        q = squareform(pdist(_ROOT3 * np.atleast_2d(x[:, th] / ell[th]).T, 'euclidean'))
        dk[th] = q * q * eQ
    return sf2 * k, sf2 * dk


def nlml_grad(hypers, x, y, mX):
    """Negative log marginal likelihood and the gradient *as the reference codes it*,
    GPR_CS2S3.py:107-141.  Components 3 and 4 are twice the true derivative;
    components >= 5 stay zero (:131, dead 6th variable of x0 at :217).  On a Cholesky failure
    returns (inf, inf-vector) (:139-140)."""
    hypers = np.asarray(hypers, dtype=float)
    ell = [np.exp(hypers[0]), np.exp(hypers[1]), np.exp(hypers[2])]
    sf2 = np.exp(hypers[3])
    sn2 = np.exp(hypers[4])
    n = len(y)
    K, dK = matern32(x, ell, sf2, want_grad=True)
    try:
        L = np.linalg.cholesky(K + np.eye(n) * sn2)
    except np.linalg.LinAlgError:
        return np.inf, np.ones(len(hypers)) * np.inf
    r = y - mX
    A = np.atleast_2d(np.linalg.solve(L.T, np.linalg.solve(L, r))).T
    nlZ = np.dot(r.T, A) / 2 + np.log(L.diagonal()).sum() + n * _LOG2PI / 2
    Qm = np.linalg.solve(L.T, np.linalg.solve(L, np.eye(n))) - np.dot(A, A.T)
    g = np.zeros(len(hypers))
    for th in range(min(len(hypers), 5)):
        if th < 3:
            g[th] = (Qm * dK[th]).sum() / 2
        elif th == 3:
            g[th] = (Qm * (2 * K)).sum() / 2
        else:
            g[th] = sn2 * np.trace(Qm)
    return float(nlZ[0]), g


def neighbours(tree, centre_xy, radius_m):
    """Observation indices within radius of a cell centre, GPR_CS2S3.py:159 (inclusive <=,
    p=2, unsorted tree order)."""
    return tree.query_ball_point(x=centre_xy, r=radius_m)


def neighbours_brute(x_train, y_train, centre_xy, radius_m):
    """Same set by direct evaluation of dx*dx+dy*dy <= r*r (sorted ascending)."""
    dx = x_train - centre_xy[0]
    dy = y_train - centre_xy[1]
    return np.nonzero(dx * dx + dy * dy <= radius_m * radius_m)[0]


def predict(inputs, outputs, mean, Xs, ell, sf2, sn2):
    """Posterior at one target, GPR_CS2S3.py:173-191.  Returns (fs, sfs2 [std-dev], lZ), or None
    when np.linalg raises LinAlgError (the reference then returns its NaN tuple, :187-191)."""
    n = len(outputs)
    mX = np.ones(n) * mean
    Kx = matern32(inputs, ell, sf2)
    Kxsx = matern32(inputs, ell, sf2, xs=Xs)
    Kxs = matern32(Xs, ell, sf2)
    try:
        L = np.linalg.cholesky(Kx + np.eye(n) * sn2)
        A = np.linalg.solve(L.T, np.linalg.solve(L, (outputs - mX)))
        lZ = -np.dot((outputs - mX).T, A) / 2 - np.log(L.diagonal()).sum() - n * _LOG2PI / 2
        v = np.linalg.solve(L, Kxsx)
        fs = mean + np.dot(Kxsx.T, A)
        with np.errstate(invalid='ignore'):
            sfs2 = np.sqrt((Kxs - np.dot(v.T, v)).diagonal())
    except np.linalg.LinAlgError:
        return None
    return float(fs[0]), float(sfs2[0]), float(lZ)


def fit(inputs, outputs, mean, x0, return_result=False):
    """Hyperparameter fit, GPR_CS2S3.py:166: scipy CG, jac=True, defaults, status ignored."""
    mX = np.ones(len(outputs)) * mean
    res = scipy.optimize.minimize(nlml_grad, x0=list(x0), args=(inputs, outputs, mX),
                                  method='CG', jac=True)
    h = np.exp(res.x)
    return (h, res) if return_result else h


class DayOracle:
    """The reference's GPR3D (GPR_CS2S3.py:143-191) over a day's arrays (a9 globals)."""

    def __init__(self, x_train, y_train, t_train, z, X, radius_km, mean, T_mid, x0):
        self.x_train, self.y_train, self.t_train, self.z = x_train, y_train, t_train, z
        self.X, self.radius_m, self.mean, self.T_mid = X, radius_km * 1000, mean, T_mid
        self.x0 = list(x0)
        self.tree = scipy.spatial.cKDTree(np.array([x_train, y_train]).T)   # :245-246

    @classmethod
    def from_day(cls, day, x0=None):
        return cls(day.x_train, day.y_train, day.t_train, day.z, day.X, day.radius_km,
                   day.mean, day.T_mid, day.x0 if x0 is None else x0)

    def cell_data(self, index, sort=False):
        ID = neighbours(self.tree, self.X[index, :], self.radius_m)
        if sort:
            ID = sorted(ID)
        inputs = np.array([self.x_train[ID], self.y_train[ID], self.t_train[ID]]).T
        outputs = self.z[ID]
        Xs = np.atleast_2d(np.array([self.X[index, 0], self.X[index, 1], self.T_mid]))
        return ID, inputs, outputs, Xs

    def gpr3d(self, index, hypers=None, sort=False, return_result=False):
        """opt=True when ``hypers`` is None (fit, :166-168), else predict with the given
        natural-unit (lx, ly, lt, sf2, sn2) (:169-172).  Returns the reference's 8-tuple."""
        _, inputs, outputs, Xs = self.cell_data(index, sort)
        res = None
        if hypers is None:
            h, res = fit(inputs, outputs, self.mean, self.x0, return_result=True)
        else:
            h = np.asarray(hypers, dtype=float)
        p = predict(inputs, outputs, self.mean, Xs, [h[0], h[1], h[2]], h[3], h[4])
        if p is None:
            out = (np.nan,) * 8
        else:
            out = (p[0], p[1], p[2], h[0], h[1], h[2], h[3], h[4])
        return (out, res) if return_result else out
