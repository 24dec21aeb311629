"""Generate tests/golden/cfg5_fit_sample.npz: the reference path (oracle restatement of GPR3D) fitted on cells of the
BASELINE.json configs[4] workload (12.5 km lattice, 500 km radius, n in the thousands; ``make_day_cfg5``), observations in
ascending index order.  16 candidate cells spread over the n range of every 64th ice cell; results are checkpointed as
they arrive (a fit at n ~ 4000 is 170+ evaluations of ~10 s each on one core), tests use whatever is marked done.

    nice -n 19 python tests/golden/make_cfg5_fit_sample.py
"""
import os, sys, time, warnings
os.environ["OPENBLAS_NUM_THREADS"] = "1"
import multiprocessing as mp
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from optimalinterpolation_b200.synthetic import make_day_cfg5   # noqa: E402

OUT = os.path.join(HERE, "cfg5_fit_sample.npz")
QUANTILES = np.linspace(0.01, 0.93, 16)
_G = {}


def _init():
    warnings.simplefilter("ignore")
    from oracle.gpr_oracle import DayOracle
    _G["o"] = DayOracle.from_day(make_day_cfg5())


def _work(c):
    t0 = time.time()
    out, res = _G["o"].gpr3d(int(c), sort=True, return_result=True)
    return int(c), np.array(out, dtype=float), int(res.nfev), int(res.status), float(res.fun), time.time() - t0


def choose_cells(d):
    from scipy.spatial import cKDTree
    sub = np.arange(0, len(d.X), 64)
    cnt = np.asarray(cKDTree(np.c_[d.x_train, d.y_train]).query_ball_point(d.X[sub], r=d.radius_km * 1000.0, return_length=True))
    o = np.argsort(cnt, kind="stable")
    pick = o[(QUANTILES * (len(o) - 1)).round().astype(int)]
    return sub[pick], cnt[pick]


def main():
    d = make_day_cfg5()
    cells, n = choose_cells(d)
    print("cells", cells, "n", n, flush=True)
    rec = {}
    with mp.get_context("fork").Pool(int(os.environ.get("OI_GOLDEN_WORKERS", 4)), initializer=_init) as pool:
        # smallest first: they finish first and the checkpoint is useful early
        for k, (c, out, nfev, status, fun, sec) in enumerate(pool.imap_unordered(_work, [int(c) for c in cells], chunksize=1)):
            rec[c] = (out, nfev, status, fun, sec)
            done = np.array([int(c) in rec for c in cells])
            nan8 = np.full(8, np.nan)
            np.savez_compressed(OUT + ".tmp.npz", cells=cells, n=n, done=done, numpy=np.__version__, scipy=__import__("scipy").__version__,
                                out=np.array([rec[int(c)][0] if int(c) in rec else nan8 for c in cells]),
                                nfev=np.array([rec[int(c)][1] if int(c) in rec else -1 for c in cells]),
                                status=np.array([rec[int(c)][2] if int(c) in rec else -1 for c in cells]),
                                fun=np.array([rec[int(c)][3] if int(c) in rec else np.nan for c in cells]),
                                seconds=np.array([rec[int(c)][4] if int(c) in rec else np.nan for c in cells]))
            os.replace(OUT + ".tmp.npz", OUT)
            print(f"{k + 1}/{len(cells)} cell {c} nfev {nfev} status {status} {sec:.0f} s", flush=True)


if __name__ == "__main__":
    main()
