"""Kernel-level benchmark: one SMLII evaluation for every cell of a stripe (no optimiser, no tail).
Prints ms and algorithmic TFLOP/s per kernel family.  OI_LIB=<path> selects an experimental build."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from optimalinterpolation_b200 import _lib
if os.environ.get("OI_LIB"):
    _lib.LIB_PATH = os.environ["OI_LIB"]
import optimalinterpolation_b200 as oi
from optimalinterpolation_b200.synthetic import make_day

stride = int(sys.argv[1]) if len(sys.argv) > 1 else 16
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
d = make_day()
cells = np.arange(0, len(d.X), stride)
h = oi.Handle(0)
h.set_observations(d.x_train, d.y_train, d.t_train, d.z); h.set_cells(d.X[cells]); cnt = h.gather_neighbours(d.radius_km * 1000.0)
if os.environ.get("OI_NRANGE"):          # OI_NRANGE=lo,hi: only cells with lo <= n <= hi
    lo, hi = map(int, os.environ["OI_NRANGE"].split(","))
    cells = cells[(cnt >= lo) & (cnt <= hi)]
    h.set_cells(d.X[cells]); cnt = h.gather_neighbours(d.radius_km * 1000.0)
    print("n range", lo, hi, "->", len(cells), "cells, mean n", cnt.mean())
hyp = np.log([2.15e5, 1.40e5, 21.0, 0.0279, 0.00346, 0.1])
walls = []
for r in range(reps):
    t0 = time.perf_counter(); f, g = h.nlml_grad(hyp, d.mean); walls.append((time.perf_counter() - t0) * 1e3)
    st = h.stats()
print("wall ms per evaluation of all cells (incl. H2D/D2H of hypers and results):", [round(w, 2) for w in walls],
      " => TFLOP/s (best)", st["flops"] / min(walls) * 1e-9, "groups", st["n_groups"])
tot = sum(st[k] for k in st if k.startswith("ms_") and k not in ("ms_total", "ms_gather", "ms_factor", "ms_persistent"))
print("cells", len(cells), "sum ms", round(tot, 3), "checksum", float(np.nansum(f)), float(np.nansum(g)))
for k in ("build", "chol", "fwd", "trtri", "alpha", "lauum", "finalize"):
    fl = st.get("flops_" + k, 0)
    if st['ms_' + k] > 0:
        print(f"  {k:9s} {st['ms_' + k]:9.3f} ms" + (f"  {fl / st['ms_' + k] * 1e-9:7.2f} TFLOP/s" if fl else ""))
print("  factor TFLOP/s", st["flops_factor"] / max(st["ms_factor"], 1e-9) * 1e-9)
