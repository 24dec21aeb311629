"""Generate tests/golden/small_day_fit.npz: the reference path (oracle restatement of GPR3D) fitted on every 4th cell of the
small synthetic day (make_small_day: n ~ 40...300), in the reference's neighbour order and in ascending index order
(see make_day_fit_sample_1k.py for why both).      python tests/golden/make_small_day_fit.py   # a few minutes"""
import os, sys, warnings
os.environ["OPENBLAS_NUM_THREADS"] = "1"
import multiprocessing as mp
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from optimalinterpolation_b200.synthetic import make_small_day   # noqa: E402
_G = {}


def _init():
    warnings.simplefilter("ignore")
    from oracle.gpr_oracle import DayOracle
    _G["o"] = DayOracle.from_day(make_small_day())


def _work(t):
    c, srt = t
    out, res = _G["o"].gpr3d(int(c), sort=bool(srt), return_result=True)
    return c, srt, np.array(out, dtype=float), int(res.nfev), int(res.status)


if __name__ == "__main__":
    d = make_small_day()
    cells = np.arange(0, len(d.X), 4)
    with mp.get_context("fork").Pool(int(os.environ.get("OI_GOLDEN_WORKERS", 3)), initializer=_init) as pool:
        rows = pool.map(_work, [(int(c), s) for c in cells for s in (0, 1)], chunksize=4)
    rec = {(r[0], r[1]): r for r in rows}
    np.savez_compressed(os.path.join(HERE, "small_day_fit.npz"), cells=cells, numpy=np.__version__, scipy=__import__("scipy").__version__,
                        out_tree=np.array([rec[(int(c), 0)][2] for c in cells]), out_sorted=np.array([rec[(int(c), 1)][2] for c in cells]),
                        nfev_tree=np.array([rec[(int(c), 0)][3] for c in cells]), nfev_sorted=np.array([rec[(int(c), 1)][3] for c in cells]),
                        status_tree=np.array([rec[(int(c), 0)][4] for c in cells]), status_sorted=np.array([rec[(int(c), 1)][4] for c in cells]))
    print("cells", len(cells))
