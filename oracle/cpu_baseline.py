"""CPU baseline: the reference's per-cell path (oracle restatement of GPR3D) on all host cores.

TEST/BENCH INFRASTRUCTURE ONLY.  Mirrors the reference's one-MPI-rank-per-core data parallelism
over cells (GPR_CS2S3.py:18-23, :250-262) with multiprocessing, one BLAS thread per worker
(mpi4py/mpirun are not installed).  A full day is ~20 CPU-hours, so a bounded sample is timed:
``cores`` cells taken evenly from the cheapest ``frac`` of the workload's n-sorted cells, and the
measured cells/s is scaled to the workload's cost mix by the n^3 cost ratio (one SMLII
evaluation is ~7 n^3 flops in the reference).
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


def choose_sample(counts, cells, n_sample, frac):
    """``n_sample`` cells evenly spaced over the cheapest ``frac`` of ``cells`` (sorted by n)."""
    cells = np.asarray(cells)
    order = cells[np.argsort(counts[cells], kind="stable")]
    top = max(n_sample, int(len(order) * frac))
    pick = np.unique(np.linspace(0, top - 1, n_sample).round().astype(int))
    return order[pick]


def run_sample(day, counts, cells, cores=None, frac=0.25, n_sample=None, x0=None):
    """Time the oracle's GPR3D on a bounded sample.  Returns a dict with cells/s scaled to the
    cost mix of ``cells`` and a description of the sample."""
    cores = cores or os.cpu_count() or 1
    n_sample = n_sample or cores
    x0 = list(day.x0 if x0 is None else x0)
    sample = choose_sample(counts, cells, n_sample, frac)
    arrays = (day.x_train, day.y_train, day.t_train, day.z, day.X, day.radius_km, day.mean, day.T_mid)
    ctx = mp.get_context("fork")
    with ctx.Pool(cores, initializer=_init, initargs=(arrays, x0)) as pool:
        pool.map(_work, [int(sample[0])] * 0)          # spin the workers up
        t0 = time.perf_counter()
        rows = pool.map(_work, [int(c) for c in sample], chunksize=1)
        wall = time.perf_counter() - t0
    n3_sample = float(np.mean(counts[sample].astype(np.float64) ** 3))
    n3_work = float(np.mean(counts[np.asarray(cells)].astype(np.float64) ** 3))
    raw = len(sample) / wall
    return dict(value=raw * n3_sample / n3_work, raw_cells_per_s=raw, wall_s=wall, cores=cores,
                n_sample=len(sample), n_min=int(counts[sample].min()), n_max=int(counts[sample].max()),
                cost_ratio=n3_sample / n3_work, nfev_mean=float(np.mean([r[2] for r in rows])),
                rows=rows)
