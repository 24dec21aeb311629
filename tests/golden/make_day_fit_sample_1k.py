"""Generate tests/golden/day_fit_sample_1k.npz: the reference path (oracle = restatement of GPR3D, bit-identical to the
reference functions, tests/test_oracle.py) fitted on 1024 cells of the FULL synthetic 25 km day, stratified over the
WHOLE n range (158 ... 1786: every 1024-quantile of the day's sorted n), each cell fitted twice:

  * ``out_tree``   -- observations in the reference's neighbour order (cKDTree.query_ball_point, GPR_CS2S3.py:159);
  * ``out_sorted`` -- the same observations in ascending index order (a permutation of the same inputs; it is the order
                      the CUDA path uses).  The reference's stopping point is not invariant to such a permutation
                      (SURVEY.md C.8), so |out_tree - out_sorted| is the reference-vs-itself floor that the parity gate
                      of tests/test_gpu_day.py is measured against.

    nice -n 19 python tests/golden/make_day_fit_sample_1k.py     # ~37 core-hours; checkpoints every 32 results

Results are checkpointed into the .npz as they arrive (cells in a seeded random order, so any prefix is itself a
stratified sample); re-running resumes from the checkpoint.
"""
import os, sys, time, warnings
os.environ["OPENBLAS_NUM_THREADS"] = "1"      # before numpy loads OpenBLAS: one BLAS thread per worker process
import multiprocessing as mp
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from optimalinterpolation_b200.synthetic import make_day   # noqa: E402

OUT = os.path.join(HERE, "day_fit_sample_1k.npz")
N_CELLS = 1024
_G = {}


def _init():
    warnings.simplefilter("ignore")
    from oracle.gpr_oracle import DayOracle
    _G["o"] = DayOracle.from_day(make_day())


def _work(task):
    c, srt = task
    t0 = time.time()
    out, res = _G["o"].gpr3d(int(c), sort=bool(srt), return_result=True)
    return c, srt, np.array(out, dtype=float), int(res.nfev), int(res.status), float(res.fun), time.time() - t0


def choose_cells(counts):
    order = np.argsort(counts, kind="stable")
    return order[np.linspace(0, len(order) - 1, N_CELLS).round().astype(int)]


def save(cells, counts, rec):
    def col(k, srt, fill, dtype=float):
        return np.array([rec.get((int(c), srt), {}).get(k, fill) for c in cells], dtype=dtype)
    nan8 = np.full(8, np.nan)
    tmp = OUT + ".tmp.npz"
    np.savez_compressed(
        tmp, cells=cells, n=counts[cells], numpy=np.__version__, scipy=__import__("scipy").__version__,
        done_tree=col("done", 0, False, bool), done_sorted=col("done", 1, False, bool),
        out_tree=np.array([rec.get((int(c), 0), {}).get("out", nan8) for c in cells]),
        out_sorted=np.array([rec.get((int(c), 1), {}).get("out", nan8) for c in cells]),
        nfev_tree=col("nfev", 0, -1, int), nfev_sorted=col("nfev", 1, -1, int),
        status_tree=col("status", 0, -1, int), status_sorted=col("status", 1, -1, int),
        fun_tree=col("fun", 0, np.nan), fun_sorted=col("fun", 1, np.nan),
        seconds_tree=col("seconds", 0, np.nan), seconds_sorted=col("seconds", 1, np.nan))
    os.replace(tmp, OUT)


def main():
    from scipy.spatial import cKDTree
    d = make_day()
    counts = np.asarray(cKDTree(np.c_[d.x_train, d.y_train]).query_ball_point(d.X, r=d.radius_km * 1000.0, return_length=True))
    cells = choose_cells(counts)
    rec = {}
    if os.path.exists(OUT):
        old = np.load(OUT)
        if np.array_equal(old["cells"], cells):
            for i, c in enumerate(cells):
                for srt, tag in ((0, "tree"), (1, "sorted")):
                    if old["done_" + tag][i]:
                        rec[(int(c), srt)] = dict(done=True, out=old["out_" + tag][i], nfev=int(old["nfev_" + tag][i]),
                                                  status=int(old["status_" + tag][i]), fun=float(old["fun_" + tag][i]),
                                                  seconds=float(old["seconds_" + tag][i]))
    perm = np.random.default_rng(1).permutation(len(cells))
    tasks = [(int(cells[i]), srt) for i in perm for srt in (0, 1) if (int(cells[i]), srt) not in rec]
    print("cells", len(cells), "n", counts[cells].min(), counts[cells].max(), "tasks to run", len(tasks), flush=True)
    workers = int(os.environ.get("OI_GOLDEN_WORKERS", os.cpu_count()))
    t0 = time.time()
    with mp.get_context("fork").Pool(workers, initializer=_init) as pool:
        for k, (c, srt, out, nfev, status, fun, sec) in enumerate(pool.imap_unordered(_work, tasks, chunksize=1)):
            rec[(c, srt)] = dict(done=True, out=out, nfev=nfev, status=status, fun=fun, seconds=sec)
            if (k + 1) % 32 == 0 or k + 1 == len(tasks):
                save(cells, counts, rec)
                print(f"{k + 1}/{len(tasks)} results, {time.time() - t0:.0f} s", flush=True)
    print("cpu seconds total", sum(r["seconds"] for r in rec.values()))


if __name__ == "__main__":
    main()
