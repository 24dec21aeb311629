"""Both engines on the same cells: NLML/gradient and fitted outputs must be bit-identical.
usage: engine_check.py [n_cells_small_day] [group_size]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import optimalinterpolation_b200 as oi
from optimalinterpolation_b200.synthetic import make_small_day

ncell = int(sys.argv[1]) if len(sys.argv) > 1 else 40
gs = int(sys.argv[2]) if len(sys.argv) > 2 else 0
d = make_small_day()
cells = np.linspace(0, len(d.X) - 1, ncell).round().astype(int)
hyp = np.log([2.15e5, 1.40e5, 21.0, 0.0279, 0.00346, 0.1])
out = {}
for eng in (0, 1):
    h = oi.Handle(0)
    h.set_observations(d.x_train, d.y_train, d.t_train, d.z); h.set_cells(d.X[cells]); h.gather_neighbours(d.radius_km * 1000.0)
    os.environ["OI_ENGINE"] = str(eng)
    f, g = h.nlml_grad(hyp, d.mean)
    p = h.make_params(d.radius_km * 1000.0, d.T_mid, d.mean, d.x0, mode=0, engine=eng, group_size=gs)
    t0 = time.time(); h.run(p); dt = time.time() - t0
    r = h.get_results()
    out[eng] = (f, g, r["out"], r["nfev"], r["status"])
    print("engine", eng, "fit s", round(dt, 3), "nfev mean", r["nfev"].mean(), "status", np.bincount(r["status"]), "n", r["n"].min(), r["n"].max())
    h.close()
a, b = out[0], out[1]
print("nlml identical", np.array_equal(a[0], b[0], equal_nan=True), "grad identical", np.array_equal(a[1], b[1], equal_nan=True))
print("fit out identical", np.array_equal(a[2], b[2], equal_nan=True), "nfev identical", np.array_equal(a[3], b[3]),
      "status identical", np.array_equal(a[4], b[4]))
if not np.array_equal(a[0], b[0], equal_nan=True):
    print("max rel nlml diff", np.nanmax(np.abs(a[0] - b[0]) / np.abs(a[0])))
