"""Per-cell fit parity report against tests/golden/day_fit_sample.npz (reference path run on the CPU)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import optimalinterpolation_b200 as oi
from optimalinterpolation_b200.synthetic import make_day
name = sys.argv[1] if len(sys.argv) > 1 else "day_fit_sample.npz"
only_outliers = len(sys.argv) > 2
g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", name))
day = make_day()
cells = g["cells"]
gd = oi.GPRDay(day.x_train, day.y_train, day.t_train, day.z, day.X[cells], day.radius_km, day.mean, day.T_mid, day.x0)
res = gd.run(opt=True)
out = res["out"]
ref = g["out_tree"] if "out_tree" in g else g["out"]
ref2 = g["out_sorted"] if "out_sorted" in g else ref
nf_ref = g["nfev_tree"] if "nfev_tree" in g else g["nfev"]
nf_ref2 = g["nfev_sorted"] if "nfev_sorted" in g else nf_ref
st_ref = g["status_tree"] if "status_tree" in g else g["status"]
print("n    nfev_gpu nfev_ref nfev_ref_sorted st_gpu st_ref  dfs_mm(gpu-ref)  dfs_mm(ref_sorted-ref)  rel_dlZ(gpu-ref)  hyp_gpu / hyp_ref")
for k in range(len(cells)):
    bad = (np.isnan(out[k, 0]) != np.isnan(ref[k, 0])) or abs(out[k, 0] - ref[k, 0]) > 1e-4 or abs(out[k, 2] - ref[k, 2]) > 1e-6 * abs(ref[k, 2])
    if only_outliers and not bad:
        continue
    print(f"{g['n'][k]:4d} {res['nfev'][k]:8d} {nf_ref[k]:8d} {nf_ref2[k]:8d} {res['status'][k]:6d} {st_ref[k]:6d} "
          f"{(out[k,0]-ref[k,0])*1e3:14.5f} {(ref2[k,0]-ref[k,0])*1e3:14.5f} {(out[k,2]-ref[k,2])/abs(ref[k,2]):12.3e}  "
          f"{np.array2string(out[k,3:], precision=4)} / {np.array2string(ref[k,3:], precision=4)}  std {out[k,1]:.5f} / {ref[k,1]:.5f}  lZ {out[k,2]:.3f} / {ref[k,2]:.3f}")
