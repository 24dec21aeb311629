"""Generate tests/golden/day_fit_sample.npz: the reference path (oracle = restatement of GPR3D, bit-identical
to the reference functions, tests/test_oracle.py) fitted on a stratified sample of cells of the FULL synthetic
25 km day (n up to ~1000), in the reference's neighbour order (tree order) and in ascending order.  The second
run measures the reference's own sensitivity to observation order (SURVEY.md C.8 "noise floor").

    python tests/golden/make_day_fit_sample.py        # ~20 min on 8 cores
"""
import os, sys, time, warnings
os.environ["OPENBLAS_NUM_THREADS"] = "1"      # before numpy loads OpenBLAS: one BLAS thread per worker process
import multiprocessing as mp
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from optimalinterpolation_b200.synthetic import make_day   # noqa: E402

_G = {}


def _init():
    warnings.simplefilter("ignore")
    from oracle.gpr_oracle import DayOracle
    _G["o"] = DayOracle.from_day(make_day())


def _work(args):
    c, srt = args
    t0 = time.time()
    out, res = _G["o"].gpr3d(int(c), sort=bool(srt), return_result=True)
    return c, srt, np.array(out, dtype=float), res.nfev, res.status, res.nit, time.time() - t0


def main():
    from scipy.spatial import cKDTree
    d = make_day()
    counts = np.asarray(cKDTree(np.c_[d.x_train, d.y_train]).query_ball_point(d.X, r=d.radius_km * 1000.0, return_length=True))
    stripe = np.arange(0, len(d.X), 16)
    order = stripe[np.argsort(counts[stripe], kind="stable")]
    order = order[counts[order] <= 1000]
    cells = order[np.linspace(0, len(order) - 1, 24).round().astype(int)]
    jobs = [(int(c), s) for s in (0, 1) for c in cells[::-1]]
    with mp.get_context("fork").Pool(os.cpu_count(), initializer=_init) as pool:
        rows = pool.map(_work, jobs, chunksize=1)
    out = {"cells": cells, "n": counts[cells], "numpy": np.__version__, "scipy": __import__("scipy").__version__}
    for srt, name in ((0, "tree"), (1, "sorted")):
        sel = {r[0]: r for r in rows if r[1] == srt}
        out[f"out_{name}"] = np.array([sel[int(c)][2] for c in cells])
        out[f"nfev_{name}"] = np.array([sel[int(c)][3] for c in cells])
        out[f"status_{name}"] = np.array([sel[int(c)][4] for c in cells])
        out[f"seconds_{name}"] = np.array([sel[int(c)][6] for c in cells])
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "day_fit_sample.npz"), **out)
    dfs = np.abs(out["out_tree"][:, 0] - out["out_sorted"][:, 0]) * 1e3
    print("cells", list(cells), "n", list(counts[cells]))
    print("reference vs re-ordered reference |dfs| mm:", np.round(dfs, 5))
    print("cpu seconds per cell:", np.round(out["seconds_tree"], 1))


if __name__ == "__main__":
    main()
