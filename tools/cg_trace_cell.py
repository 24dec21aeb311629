"""Evaluation-by-evaluation comparison of the device's own optimiser run (oi_debug_trace) with the host build of the same
state machine driven by the GPU objective (oi_nlml_grad), for one cell of the day alone and inside a batch.  (needs a GPU)
usage: python tools/cg_trace_cell.py <day cell index> [<day cell index> ...]"""
import os, sys, warnings
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import optimalinterpolation_b200 as oi
from optimalinterpolation_b200.synthetic import make_day
from cg_driver import minimize_cg
warnings.simplefilter("ignore")
np.set_printoptions(linewidth=220, precision=8)
day = make_day()
f = np.load(os.path.join(ROOT, "tests", "golden", "day_fit_sample_1k.npz"))
use = f["done_tree"] & f["done_sorted"]
batch = f["cells"][use]
h = oi.Handle(0)
h.set_observations(day.x_train, day.y_train, day.t_train, day.z)
R = day.radius_km * 1000.0
out = {}
for cell in [int(a) for a in sys.argv[1:]]:
    # (1) alone
    h.set_cells(day.X[cell][None, :]); h.gather_neighbours(R)
    h.debug_trace(0, 8192)
    h.run(h.make_params(R, day.T_mid, day.mean, day.x0, mode=0))
    r1 = h.get_results(); t1 = h.get_debug_trace()
    # host state machine on the API objective
    evals = []

    def fun(x):
        nlz, g = h.nlml_grad(np.asarray(x, dtype=float)[None, :], day.mean)
        evals.append(np.r_[np.exp(x[:5]), nlz[0], g[0]])
        return float(nlz[0]), g[0].copy()
    h.debug_trace(-1, 0)
    H = minimize_cg(fun, day.x0)
    tH = np.array(evals)
    # (2) inside the batch
    pos = int(np.nonzero(batch == cell)[0][0])
    h.set_cells(day.X[batch]); h.gather_neighbours(R)
    h.debug_trace(pos, 8192)
    h.run(h.make_params(R, day.T_mid, day.mean, day.x0, mode=0))
    r2 = h.get_results(); t2 = h.get_debug_trace()
    h.debug_trace(-1, 0)
    print(f"cell {cell}: alone nfev {r1['nfev'][0]} st {r1['status'][0]} lZ {r1['out'][0, 2]:.6f} | batch nfev {r2['nfev'][pos]} st {r2['status'][pos]} lZ {r2['out'][pos, 2]:.6f} | "
          f"host nfev {H['nfev']} st {H['status']} | trace rows alone {len(t1)} batch {len(t2)} host {len(tH)}")
    for name, t in (("alone", t1), ("batch", t2)):
        m = min(len(t), len(tH))
        dh = np.abs(t[:m, :5] / tH[:m, :5] - 1).max(axis=1)
        k = next((i for i in range(m) if not np.array_equal(t[i, :5], tH[i, :5])), None)
        print(f"  {name}: first evaluation whose hyperparameters differ from the host run: {k}")
        if k is not None:
            for i in range(max(0, k - 2), min(m, k + 3)):
                print(f"   eval {i}: rel diff hyp {dh[i]:.2e} | device f {t[i, 5]!r} host f {tH[i, 5]!r} | device g {t[i, 6:11]} host g {tH[i, 6:11]}")
        big = next((i for i in range(m) if dh[i] > 1e-6), None)
        print(f"  {name}: first evaluation with hyperparameters off by > 1e-6: {big}")
        if big is not None:
            for i in range(max(0, big - 3), min(m, big + 2)):
                # re-evaluate the device's own point through the API
                nlz, g = h.nlml_grad(np.r_[np.log(t[i, :5]), np.log(0.1)][None, :], day.mean) if name == "alone" and False else (None, None)
                print(f"   eval {i}: device hyp {np.log(t[i, :5])} f {t[i, 5]!r} g {t[i, 6:11]}\n            host   hyp {np.log(tH[i, :5])} f {tH[i, 5]!r} g {tH[i, 6:11]}")
    out[f"alone_{cell}"] = t1; out[f"batch_{cell}"] = t2; out[f"host_{cell}"] = tH
np.savez_compressed(os.path.join(ROOT, "gpurun_out", "cg_trace_cell.npz"), **out)
