"""Run a stripe of the synthetic pan-Arctic day on the GPU and print timing/stats."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from optimalinterpolation_b200 import _lib
if os.environ.get("OI_LIB"):
    _lib.LIB_PATH = os.environ["OI_LIB"]
import optimalinterpolation_b200 as oi
from optimalinterpolation_b200.synthetic import make_day

stride = int(sys.argv[1]) if len(sys.argv) > 1 else 8
max_active = int(sys.argv[2]) if len(sys.argv) > 2 else 0
d = make_day()
cells = np.arange(0, len(d.X), stride)
g = oi.GPRDay(d.x_train, d.y_train, d.t_train, d.z, d.X[cells], d.radius_km, d.mean, d.T_mid, d.x0)
t0 = time.time(); res = g.run(opt=True, max_active=max_active); dt = time.time() - t0
st = g.handle.stats()
print("cells", len(cells), "wall s", dt, "cells/s", len(cells) / dt)
print("TFLOP/s overall", st["flops"] / st["ms_total"] * 1e-9, "factor kernels", st["flops_factor"] / st["ms_factor"] * 1e-9)
for k in ("chol", "trtri", "lauum"):
    if st["ms_" + k] > 0:
        print(k, "TFLOP/s", st["flops_" + k] / st["ms_" + k] * 1e-9, "ms", st["ms_" + k])
print("graph captures", st["n_graph_captures"], "graph launches", st["n_graph_launches"], "ms_graph", round(st["ms_graph"], 1))
print("express cells", st["n_express_cells"], "iterations", st["n_iterations"], "launches", st["n_launches"])
print("nfev mean", res["nfev"].mean(), "max", res["nfev"].max(), "status hist", np.bincount(res["status"]))
print("n mean", res["n"].mean(), "out finite frac", np.isfinite(res["out"][:, 0]).mean())
np.save("gpurun_out/day_stripe_out.npy", res["out"])
