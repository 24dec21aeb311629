"""Host-side mirror of the reference's per-cell GP interface, calling the CUDA library.

The reference's seam is ``GPR3D(index, opt=True)`` (/root/reference/2021_paper_production/
GPR_CS2S3.py:143-191), called per cell from the rank loops (:258-261, :317-319) and reading module
globals (:201-246).  ``GPRDay`` takes those globals as constructor arguments (same names) and
evaluates ALL cells in one call; ``GPRDay.GPR3D(index)`` then returns the reference's tuple for
one cell so existing post-processing (:264-297) keeps working.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib

STATUS = {0: "ok", 1: "maxiter", 2: "line search failed (scipy status 2)", 3: "cholesky failed",
          4: "no observations", 5: "nan"}


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _ptr(a):
    return None if a is None else a.ctypes.data


class OIError(RuntimeError):
    pass


class Handle:
    """Thin object wrapper over the opaque ``oi_handle`` (one per GPU)."""

    def __init__(self, device: int = 0):
        self._L = _lib.load()
        self._h = C.c_void_p()
        self._check(self._L.oi_create(int(device), C.byref(self._h)))
        self.n_cells = 0
        self.n_obs = 0

    def _check(self, rc):
        if rc != 0:
            raise OIError(f"liboi_b200 error {rc}: {self._L.oi_last_error().decode()}")

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._L.oi_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream_ptr):
        """Launch on a caller-owned stream (``torch.cuda.Stream().cuda_stream``).  0 / None restores the handle's own
        stream: torch's legacy default stream has the handle 0 and therefore cannot be selected."""
        self._check(self._L.oi_set_stream(self._h, C.c_void_p(cuda_stream_ptr) if cuda_stream_ptr else None))

    def set_observations(self, x, y, t, z):
        x, y, t, z = _f64(x), _f64(y), _f64(t), _f64(z)
        if not (x.shape == y.shape == t.shape == z.shape and x.ndim == 1):
            raise ValueError("x, y, t, z must be 1-D arrays of one length")
        self._check(self._L.oi_set_observations(self._h, _ptr(x), _ptr(y), _ptr(t), _ptr(z), x.size))
        self.n_obs = x.size

    def set_time_window(self, t_lo: float = -np.inf, t_hi: float = np.inf):
        """Only observations with t_lo <= t <= t_hi take part in the gather; t then counts from t_lo (the reference's
        ``obs[:, :, :, day:day+T]`` window, GPR_CS2S3.py:213).  No arguments: window off."""
        self._check(self._L.oi_set_time_window(self._h, float(t_lo), float(t_hi)))

    def set_cells(self, X):
        X = _f64(X)
        if X.ndim != 2 or X.shape[1] != 2:
            raise ValueError("X must be (n_cells, 2)")
        self._check(self._L.oi_set_cells(self._h, _ptr(X), X.shape[0]))
        self.n_cells = X.shape[0]

    def gather_neighbours(self, radius_m: float) -> np.ndarray:
        counts = np.zeros(self.n_cells, dtype=np.int32)
        self._check(self._L.oi_gather_neighbours(self._h, float(radius_m), _ptr(counts)))
        self._counts = counts
        return counts

    def get_neighbours(self):
        offsets = np.zeros(self.n_cells + 1, dtype=np.int64)
        self._check(self._L.oi_get_neighbours(self._h, _ptr(offsets), None))
        indices = np.zeros(max(int(offsets[-1]), 1), dtype=np.int32)
        self._check(self._L.oi_get_neighbours(self._h, None, _ptr(indices)))
        return offsets, indices[:int(offsets[-1])]

    def nlml_grad(self, hypers, prior_mean: float, grad_convention: int = 0):
        """SMLII for every cell: hypers (n_cells, n_hyp) LOG hyperparameters -> (nlZ, grad)."""
        hypers = _f64(hypers)
        if hypers.ndim == 1:
            hypers = np.tile(hypers, (self.n_cells, 1))
        n_hyp = hypers.shape[1]
        nlz = np.zeros(self.n_cells)
        grad = np.zeros((self.n_cells, n_hyp))
        self._check(self._L.oi_nlml_grad(self._h, _ptr(hypers), n_hyp, float(prior_mean), int(grad_convention),
                                         _ptr(nlz), _ptr(grad)))
        return nlz, grad

    def make_params(self, radius_m, t_pred, prior_mean, x0=None, mode=0, grad_convention=0, maxiter=0,
                    gtol=0.0, scratch_gib=0.0, max_active=0, n_groups=0, engine=0, group_size=0,
                    evals_per_launch=0, optimiser=0):
        p = _lib.OiParams()
        p.radius_m, p.t_pred, p.prior_mean = float(radius_m), float(t_pred), float(prior_mean)
        x0 = [0.0] * 5 if x0 is None else list(x0)
        p.n_hyp = len(x0)
        for i, v in enumerate(x0):
            p.x0[i] = float(v)
        p.mode, p.grad_convention, p.maxiter = int(mode), int(grad_convention), int(maxiter)
        p.gtol, p.scratch_gib, p.max_active = float(gtol), float(scratch_gib), int(max_active)
        p.n_groups, p.engine, p.group_size, p.evals_per_launch = int(n_groups), int(engine), int(group_size), int(evals_per_launch)
        p.optimiser = int(optimiser)     # 0 = scipy-CG restatement (parity mode), 1 = exact-gradient L-BFGS (fast mode)
        return p

    def run(self, params, hypers_in=None):
        hin = None if hypers_in is None else _f64(hypers_in)
        if hin is not None and hin.shape != (self.n_cells, 5):
            raise ValueError("hypers_in must be (n_cells, 5)")
        self._check(self._L.oi_run(self._h, C.byref(params), _ptr(hin)))

    def get_results(self):
        out = np.zeros((self.n_cells, 8))
        n = np.zeros(self.n_cells, dtype=np.int32)
        nfev = np.zeros(self.n_cells, dtype=np.int32)
        status = np.zeros(self.n_cells, dtype=np.int32)
        self._check(self._L.oi_get_results(self._h, _ptr(out), _ptr(n), _ptr(nfev), _ptr(status)))
        return dict(out=out, n=n, nfev=nfev, status=status)

    def set_shared_queue(self, name):
        """Share one cost-sorted work list with the other GPU processes of the box (same fresh name on every rank, before
        the first run; None detaches).  See include/oi_b200.h."""
        self._check(self._L.oi_set_shared_queue(self._h, None if name is None else name.encode()))

    def unlink_shared_queue(self, name):
        self._L.oi_unlink_shared_queue(name.encode())

    def get_owned(self) -> np.ndarray:
        """bool mask of the cells this handle computed in the last run."""
        m = np.zeros(self.n_cells, dtype=np.uint8)
        self._check(self._L.oi_get_owned(self._h, _ptr(m)))
        return m.astype(bool)

    def debug_trace(self, cell: int, capacity: int = 4096):
        """Record every objective evaluation of one cell's optimiser during the next fits (diagnostic)."""
        self._dbg_cap = int(capacity)
        self._check(self._L.oi_debug_trace(self._h, int(cell), int(capacity)))

    def get_debug_trace(self) -> np.ndarray:
        rows = np.zeros((self._dbg_cap, 12)); n = C.c_int32(0)
        self._check(self._L.oi_get_debug_trace(self._h, _ptr(rows), C.addressof(n)))
        return rows[:n.value].copy()

    def stats(self) -> dict:
        s = _lib.OiStats()
        self._check(self._L.oi_get_stats(self._h, C.byref(s)))
        return {k: (list(getattr(s, k)) if k == "cycles_phase" else getattr(s, k)) for k, _ in s._fields_}

    def gpr_day(self, x, y, t, z, X, params, hypers_in=None):
        """The whole day through the single ABI call (host buffers in, host buffers out)."""
        x, y, t, z, X = _f64(x), _f64(y), _f64(t), _f64(z), _f64(X)
        hin = None if hypers_in is None else _f64(hypers_in)
        nc = X.shape[0]
        out = np.zeros((nc, 8)); n = np.zeros(nc, np.int32); nfev = np.zeros(nc, np.int32); status = np.zeros(nc, np.int32)
        self._check(self._L.oi_gpr_day(self._h, _ptr(x), _ptr(y), _ptr(t), _ptr(z), x.size, _ptr(X), nc,
                                       C.byref(params), _ptr(hin), _ptr(out), _ptr(n), _ptr(nfev), _ptr(status)))
        self.n_cells, self.n_obs = nc, x.size
        return dict(out=out, n=n, nfev=nfev, status=status)


class GPRDay:
    """The reference's day globals (GPR_CS2S3.py:201-217, :238-246) + its GPR3D, batched on the GPU.

    Parameters carry the reference's names: ``x_train, y_train, t_train, z`` (flattened
    observations), ``X`` (ice-cell coordinates), ``radius`` (km), ``mean`` (prior mean), ``T_mid``
    (prediction day index), ``x0`` (initial log hyperparameters, 5 or 6 entries).
    """

    def __init__(self, x_train, y_train, t_train, z, X, radius, mean, T_mid, x0, device: int = 0,
                 grad_convention: int = 0, handle: Handle | None = None):
        self.x_train, self.y_train, self.t_train, self.z = map(_f64, (x_train, y_train, t_train, z))
        self.X = _f64(X)
        self.radius, self.mean, self.T_mid, self.x0 = float(radius), float(mean), float(T_mid), list(x0)
        self.grad_convention = grad_convention
        self.handle = handle or Handle(device)
        self._results = None
        self._results_smth = None

    @classmethod
    def from_day(cls, day, **kw):
        return cls(day.x_train, day.y_train, day.t_train, day.z, day.X, day.radius_km, day.mean, day.T_mid,
                   day.x0, **kw)

    def _params(self, mode, **kw):
        kw.setdefault("grad_convention", self.grad_convention)
        return self.handle.make_params(self.radius * 1000.0, self.T_mid, self.mean, self.x0, mode=mode, **kw)

    def run(self, opt: bool = True, ellXs=None, sf2xs=None, sn2xs=None, **kw):
        """All cells at once.  opt=True: fit + predict (pass 1, GPR_CS2S3.py:258-262).  opt=False:
        predict with the given per-cell hyperparameters (pass 2, :311-320; ``ellXs`` (n_cells,3),
        ``sf2xs``, ``sn2xs`` as at :313-315)."""
        hin = None
        if not opt:
            hin = np.column_stack([_f64(ellXs), _f64(sf2xs), _f64(sn2xs)])
        res = self.handle.gpr_day(self.x_train, self.y_train, self.t_train, self.z, self.X,
                                  self._params(0 if opt else 1, **kw), hin)
        if opt:
            self._results = res
        else:
            self._results_smth = res
        return res

    def GPR3D(self, index: int, opt: bool = True):
        """The reference's per-cell return value (GPR_CS2S3.py:184-191) from the batched results."""
        res = self._results if opt else self._results_smth
        if res is None:
            raise OIError("call run() first")
        o = res["out"][index]
        return tuple(o) if opt else (o[0], o[1])
