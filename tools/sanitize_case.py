"""A small case that touches every kernel (gather, pack, build, Cholesky update/panel/row scaling, substitution, inverse,
alpha, trace, finalize with both optimisers, predict) for compute-sanitizer:
    compute-sanitizer --tool memcheck python tools/sanitize_case.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import optimalinterpolation_b200 as oi
from optimalinterpolation_b200.synthetic import make_small_day
d = make_small_day()
cells = np.array([0, 100, 222, 300, 450, 600])          # n from a few dozen to ~300: 1 to 5 blocks, ragged edges
h = oi.Handle(0)
h.set_observations(d.x_train, d.y_train, d.t_train, d.z); h.set_cells(d.X[cells]); print("n", h.gather_neighbours(d.radius_km * 1000.0))
hyp = np.log([2.15e5, 1.40e5, 21.0, 0.0279, 0.00346, 0.1])
f, g = h.nlml_grad(hyp, d.mean); print("nlml", f)
for opt, conv in ((0, 0), (1, 1)):
    h.run(h.make_params(d.radius_km * 1000.0, d.T_mid, d.mean, d.x0, mode=0, maxiter=4, optimiser=opt, grad_convention=conv))
    r = h.get_results(); print("optimiser", opt, "fs", r["out"][:, 0], "nfev", r["nfev"])
h.run(h.make_params(d.radius_km * 1000.0, d.T_mid, d.mean, d.x0, mode=1), np.tile(np.exp(hyp[:5]), (len(cells), 1)))
print("predict", h.get_results()["out"][:, :2].ravel())
h.close()
print("sanitize case done")
