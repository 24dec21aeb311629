"""Per-cell comparison of the device fit against the CPU oracle (scipy CG on the numpy restatement)."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import optimalinterpolation_b200 as oi
from optimalinterpolation_b200.synthetic import make_small_day, make_day
from oracle.gpr_oracle import DayOracle

which = sys.argv[1] if len(sys.argv) > 1 else "small"
stride = int(sys.argv[2]) if len(sys.argv) > 2 else 12
d = make_small_day() if which == "small" else make_day()
o = DayOracle.from_day(d)
cells = np.arange(0, len(d.X), stride)
g = oi.GPRDay(d.x_train, d.y_train, d.t_train, d.z, d.X[cells], d.radius_km, d.mean, d.T_mid, d.x0)
t0 = time.time(); res = g.run(opt=True); print("gpu s", time.time() - t0, g.handle.stats())
rows = []
for k, c in enumerate(cells):
    t0 = time.time()
    ref, r = o.gpr3d(int(c), sort=True, return_result=True)
    got = res["out"][k]
    rows.append(dict(cell=int(c), n=int(res["n"][k]), ref=list(map(float, ref)), got=list(map(float, got)),
                     ref_nfev=int(r.nfev), ref_status=int(r.status), ref_nit=int(r.nit),
                     nfev=int(res["nfev"][k]), status=int(res["status"][k]), cpu_s=time.time() - t0))
    print(rows[-1]["cell"], rows[-1]["n"], "dfs_mm %.4g" % (abs(got[0] - ref[0]) * 1e3), "lZ ref %.6f got %.6f" % (ref[2], got[2]),
          "nfev", r.nfev, res["nfev"][k], "status", r.status, res["status"][k], flush=True)
json.dump(rows, open(f"gpurun_out/fit_compare_{which}.json", "w"))
