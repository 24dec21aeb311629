"""CPU baseline: the reference's per-cell path (oracle restatement of GPR3D) on all host cores.

TEST/BENCH INFRASTRUCTURE ONLY.  Mirrors the reference's one-MPI-rank-per-core data parallelism
over cells (GPR_CS2S3.py:18-23, :250-262) with multiprocessing, one BLAS thread per worker
(mpi4py/mpirun are not installed).  A full day is hundreds of CPU-hours, so a bounded sample of whole fits is timed
(``run_boxed_fits`` below; round 1's power-law model of single evaluations was replaced by it).
"""
from __future__ import annotations

import multiprocessing as mp
import os
import time

import numpy as np

_G = {}


def _init(day_arrays, x0):
    os.environ["OPENBLAS_NUM_THREADS"] = "1"
    os.environ["OMP_NUM_THREADS"] = "1"
    try:
        from threadpoolctl import threadpool_limits
        _G["lim"] = threadpool_limits(1)
    except Exception:
        pass
    import warnings
    warnings.simplefilter("ignore")
    from oracle.gpr_oracle import DayOracle
    x, y, t, z, X, radius_km, mean, T_mid = day_arrays
    _G["o"] = DayOracle(x, y, t, z, X, radius_km, mean, T_mid, x0)


def _work(index):
    t0 = time.perf_counter()
    out, res = _G["o"].gpr3d(int(index), return_result=True)
    return int(index), tuple(float(v) for v in out), int(res.nfev), time.perf_counter() - t0


# ----------------------------------------------------------------------------------------------------------
# Measured baseline (round 2): whole GPR3D fits of the reference path, time-boxed.
#
# The sample cells come from tests/golden/day_fit_sample_1k.npz: 1024 cells at the 1024-quantiles of the day's n
# distribution, for which the reference path's evaluation count (scipy's nfev; the optimiser is deterministic) is
# recorded.  Every sampled cell runs its REAL fit (scipy CG on the reference objective from x0, GPR_CS2S3.py:166)
# on one core with all other cores busy with other cells.  A fit that finishes inside the time box is measured whole
# (fit + prediction); a fit that does not is stopped and its cost is (its own measured seconds per evaluation) x
# (its recorded nfev + 1/3 for the prediction).  No power law, no evaluation count borrowed from other cells.
# ----------------------------------------------------------------------------------------------------------
FIXTURE = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "day_fit_sample_1k.npz")


class _Boxed(Exception):
    pass


def _boxed_work(args):
    index, box_s = args
    import scipy.optimize
    from oracle.gpr_oracle import nlml_grad, predict
    o = _G["o"]
    t0 = time.perf_counter()
    _, inputs, outputs, Xs = o.cell_data(int(index))
    mX = np.ones(len(outputs)) * o.mean
    calls = [0]

    def objective(h, *a):
        if calls[0] >= 3 and time.perf_counter() - t0 > box_s:
            raise _Boxed()
        calls[0] += 1
        return nlml_grad(h, inputs, outputs, mX)

    finished = True
    try:
        res = scipy.optimize.minimize(objective, x0=list(o.x0), method='CG', jac=True)
        h = np.exp(res.x)
        predict(inputs, outputs, o.mean, Xs, [h[0], h[1], h[2]], h[3], h[4])
    except _Boxed:
        finished = False
    return int(index), len(outputs), calls[0], time.perf_counter() - t0, finished


def load_fixture():
    f = np.load(FIXTURE)
    ok = f["done_tree"]
    return dict(cells=f["cells"][ok], n=f["n"][ok], nfev=f["nfev_tree"][ok], seconds=f["seconds_tree"][ok])


def run_boxed_fits(day, step_index=0, cores=None, cells_per_core=1, box_s=6.0, x0=None, fixture=None):
    """One bounded sample of whole fits: cores*cells_per_core cells at evenly spaced quantiles of the fixture (a different
    offset every ``step_index``), one process per core.  Returns per-cell rows (index, n, nfev_done, seconds, finished,
    nfev_ref, cost_seconds)."""
    cores = cores or os.cpu_count() or 1
    fx = fixture or load_fixture()
    m = len(fx["cells"])
    k = min(cores * cells_per_core, m)
    stride = m / k
    pos = ((np.arange(k) + ((step_index * 0.6180339887) % 1.0)) * stride).astype(int) % m
    pos = np.unique(pos)
    x0 = list(day.x0 if x0 is None else x0)
    arrays = (day.x_train, day.y_train, day.t_train, day.z, day.X, day.radius_km, day.mean, day.T_mid)
    t0 = time.perf_counter()
    with mp.get_context("fork").Pool(cores, initializer=_init, initargs=(arrays, x0)) as pool:
        # largest cells first so that the longest jobs do not start last
        raw = pool.map(_boxed_work, [(int(fx["cells"][p]), box_s) for p in pos[::-1]], chunksize=1)
    wall = time.perf_counter() - t0
    ref = {int(fx["cells"][p]): int(fx["nfev"][p]) for p in pos}
    rows = []
    for index, n, done, sec, finished in raw:
        cost = sec if finished else sec / done * (ref[index] + 1.0 / 3.0)
        rows.append((index, n, done, sec, finished, ref[index], cost))
    return dict(rows=rows, wall_s=wall, cores=cores, box_s=box_s)


def summarise_boxed(all_rows, cores, fixture=None):
    """cells/s of the reference path on ``cores`` cores from pooled time-boxed fits: cores / mean(cost per cell).  The
    sample is a quantile sample of the day's n distribution, and the bench's stripes have that distribution too."""
    cost = np.array([r[6] for r in all_rows])
    n = np.array([r[1] for r in all_rows], dtype=float)
    per_eval = np.array([r[3] / max(r[2], 1) for r in all_rows])
    mean_cost = float(cost.mean())
    out = dict(value=cores / mean_cost, mean_cost_s=mean_cost, sem_rel=float(cost.std(ddof=1) / np.sqrt(len(cost)) / mean_cost) if len(cost) > 1 else None,
               n_sample=len(cost), n_finished=int(sum(1 for r in all_rows if r[4])), n_min=int(n.min()), n_max=int(n.max()),
               nfev_mean=float(np.mean([r[5] for r in all_rows])), evals_timed=int(sum(r[2] for r in all_rows)))
    fx = fixture or load_fixture()
    if len(cost) >= 4:
        # cross-check with less sampling noise: the measured seconds/evaluation as a function of n (log-log interpolation of
        # the sampled cells) applied to ALL fixture cells' recorded evaluation counts
        o = np.argsort(n)
        t_all = np.exp(np.interp(np.log(fx["n"].astype(float)), np.log(n[o]), np.log(per_eval[o])))
        out["value_all_fixture_cells"] = float(cores / np.mean((fx["nfev"] + 1.0 / 3.0) * t_all))
        out["fixture_cells"] = int(len(fx["cells"]))
    return out
