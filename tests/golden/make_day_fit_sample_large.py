"""Generate tests/golden/day_fit_sample_large.npz: the reference path (oracle = restatement of GPR3D, bit-identical to
the reference functions, tests/test_oracle.py) fitted on 160 cells of the FULL synthetic 25 km day, stratified by n
(n <= 1100), in the reference's neighbour order.  Used by tests/test_gpu_day.py to measure the fraction of cells whose
fitted freeboard agrees with the reference within 1 mm (BASELINE.json north_star tolerance).

    python tests/golden/make_day_fit_sample_large.py        # ~15 min on 8 cores
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


def _work(c):
    t0 = time.time()
    out, res = _G["o"].gpr3d(int(c), return_result=True)
    return c, np.array(out, dtype=float), res.nfev, res.status, time.time() - t0


def main():
    from scipy.spatial import cKDTree
    d = make_day()
    counts = np.asarray(cKDTree(np.c_[d.x_train, d.y_train]).query_ball_point(d.X, r=d.radius_km * 1000.0, return_length=True))
    pool_cells = np.arange(5, len(d.X), 16)                 # disjoint from the cells of day_fit_sample.npz (stride 16 from 0)
    order = pool_cells[np.argsort(counts[pool_cells], kind="stable")]
    order = order[counts[order] <= 1100]
    cells = order[np.linspace(0, len(order) - 1, 160).round().astype(int)]
    with mp.get_context("fork").Pool(os.cpu_count(), initializer=_init) as pool:
        rows = pool.map(_work, [int(c) for c in cells[::-1]], chunksize=1)
    sel = {r[0]: r for r in rows}
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "day_fit_sample_large.npz"),
                        cells=cells, n=counts[cells], numpy=np.__version__, scipy=__import__("scipy").__version__,
                        out=np.array([sel[int(c)][1] for c in cells]), nfev=np.array([sel[int(c)][2] for c in cells]),
                        status=np.array([sel[int(c)][3] for c in cells]), seconds=np.array([sel[int(c)][4] for c in cells]))
    print("cells", len(cells), "n", counts[cells].min(), counts[cells].max(), "cpu seconds total", sum(r[4] for r in rows))


if __name__ == "__main__":
    main()
