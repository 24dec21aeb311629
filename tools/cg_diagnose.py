"""Why does a cell's fit on the GPU end elsewhere than the reference's?  (needs a GPU)

For every cell of tests/golden/day_fit_sample_1k.npz whose GPU fit misses the parity gate against the reference (tree order),
and for a few that meet it, three optimiser runs are compared:
  D  the device's own CG (the product path)
  S  scipy.optimize.minimize(method='CG') on the host, driven by the GPU objective (oi_nlml_grad on that one cell)
  H  the host build of the device's state machine (tests/cg_driver.py), driven by the same GPU objective
S == H  => the restatement follows scipy on this objective;  H == D => the device runs the restatement faithfully;
S != reference => the difference comes from the objective's round-off, not from the optimiser.
The evaluation points of S are saved so that the CPU objective can be recomputed along them offline.

usage: python tools/cg_diagnose.py [max_cells]   -> gpurun_out/cg_diagnose.npz + printed table
"""
import os, sys, warnings
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import scipy.optimize
import optimalinterpolation_b200 as oi
from optimalinterpolation_b200.synthetic import make_day
from cg_driver import minimize_cg

warnings.simplefilter("ignore")
max_cells = int(sys.argv[1]) if len(sys.argv) > 1 else 12
f = np.load(os.path.join(ROOT, "tests", "golden", "day_fit_sample_1k.npz"))
use = f["done_tree"] & f["done_sorted"]
cells, ref, srt = f["cells"][use], f["out_tree"][use], f["out_sorted"][use]
day = make_day()
gd = oi.GPRDay(day.x_train, day.y_train, day.t_train, day.z, day.X[cells], day.radius_km, day.mean, day.T_mid, day.x0)
res = gd.run(opt=True)
out = res["out"]
np.savez_compressed(os.path.join(ROOT, "gpurun_out", "parity_1k_gpu.npz"), cells=cells, out=out, nfev=res["nfev"], status=res["status"], n=res["n"])


def miss(a, b):
    na, nb = np.isnan(a[:, 0]), np.isnan(b[:, 0])
    both = ~na & ~nb
    bad = na != nb
    bad[both] = (np.abs(a[both, 0] - b[both, 0]) * 1e3 > 1.0) | ((a[both, 2] - b[both, 2]) / np.abs(b[both, 2]) < -1e-6)
    return bad


bad = miss(out, ref)
print(f"{len(cells)} cells, GPU misses vs tree: {bad.sum()}, vs sorted: {miss(out, srt).sum()}, reference sorted vs tree: {miss(srt, ref).sum()}")
pick = list(np.nonzero(bad)[0][:max_cells]) + list(np.nonzero(~bad)[0][:3])
h = gd.handle
rows, traces = [], {}
for k in pick:
    h.set_cells(day.X[cells[k]][None, :]); h.gather_neighbours(day.radius_km * 1000.0)

    def fun(x):
        nlz, g = h.nlml_grad(np.asarray(x, dtype=float)[None, :], day.mean)
        return float(nlz[0]), g[0].copy()
    xs, fs, gs = [], [], []

    def rec(x):
        xs.append(np.array(x)); r = fun(x); fs.append(r[0]); gs.append(r[1]); return r
    S = scipy.optimize.minimize(rec, list(day.x0), jac=True, method="CG")
    H = minimize_cg(fun, day.x0)
    traces[int(cells[k])] = np.column_stack([np.array(xs), np.array(fs), np.array(gs)])     # x (6) | f | g (6)
    row = dict(cell=int(cells[k]), n=int(res["n"][k]), missed=bool(bad[k]),
               D_nfev=int(res["nfev"][k]), D_status=int(res["status"][k]), D_fs=float(out[k, 0]), D_lZ=float(out[k, 2]),
               S_nfev=int(S.nfev), S_status=int(S.status), S_x=S.x, H_nfev=int(H["nfev"]), H_status=int(H["status"]), H_x=H["x"],
               D_hyp=out[k, 3:8], ref_fs=float(ref[k, 0]), ref_lZ=float(ref[k, 2]), ref_nfev=int(f["nfev_tree"][use][k]),
               ref_status=int(f["status_tree"][use][k]), srt_fs=float(srt[k, 0]), srt_lZ=float(srt[k, 2]), srt_nfev=int(f["nfev_sorted"][use][k]),
               srt_status=int(f["status_sorted"][use][k]))
    rows.append(row)
    SH = np.array_equal(S.x, H["x"], equal_nan=True) and S.nfev == H["nfev"]
    HD = H["nfev"] == res["nfev"][k] and np.allclose(np.exp(H["x"][:5]), out[k, 3:8], rtol=1e-12, equal_nan=True)
    print(f"cell {row['cell']:6d} n {row['n']:5d} {'MISS' if bad[k] else 'ok  '} | D nfev {row['D_nfev']:5d} st {row['D_status']} fs {row['D_fs']:.6f} lZ {row['D_lZ']:.6f} | "
          f"S nfev {S.nfev:5d} st {S.status} | H nfev {H['nfev']:5d} st {H['status']} | S==H {SH} H==D {HD} | "
          f"ref nfev {row['ref_nfev']:5d} st {row['ref_status']} fs {row['ref_fs']:.6f} lZ {row['ref_lZ']:.6f} | sorted nfev {row['srt_nfev']:5d} st {row['srt_status']} fs {row['srt_fs']:.6f} lZ {row['srt_lZ']:.6f}")
np.savez_compressed(os.path.join(ROOT, "gpurun_out", "cg_diagnose.npz"), rows=np.array(rows, dtype=object),
                    **{f"trace_{c}": t for c, t in traces.items()})
