"""Scheduler timeline of the bench step under a setting of the OI_* switches (OI_TRACE): per time bucket the active cells
per stream group, the mean block count and the algorithmic TFLOP/s.  (needs a GPU)
usage: python tools/trace_run.py "NAME=VAL,NAME=VAL" ..."""
import csv, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import numpy as np
import optimalinterpolation_b200 as oi
from optimalinterpolation_b200.synthetic import make_day
d = make_day(); nc = len(d.X)
cells = np.sort(np.concatenate([np.arange(s, nc, 16) for s in range(2)]))
h = oi.Handle(0)
h.set_observations(d.x_train, d.y_train, d.t_train, d.z); h.set_cells(d.X[cells]); h.gather_neighbours(d.radius_km * 1000.0)
p = h.make_params(d.radius_km * 1000.0, d.T_mid, d.mean, d.x0, mode=0)
h.run(p)
out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
for k, spec in enumerate(sys.argv[1:]):
    env = dict(kv.split("=") for kv in spec.split(",") if kv)
    path = os.path.join(out, f"trace_{k}.csv")
    if os.path.exists(path):
        os.remove(path)
    os.environ.update(env); os.environ["OI_TRACE"] = path
    h.run(p)
    for name in list(env) + ["OI_TRACE"]:
        os.environ.pop(name, None)
    print(f"== {spec or 'defaults'}: device ms {h.stats()['ms_total']:.1f}")
    rows = [r for r in csv.DictReader(open(path)) if r["iter"] != "iter"]
    t = np.array([float(r["t_ms"]) for r in rows]); A = np.array([int(r["active"]) for r in rows]); g = np.array([int(r["group"]) for r in rows])
    N = np.array([int(r["Nmax"]) for r in rows]); fl = np.array([float(r["flops_factor"]) for r in rows])
    G = g.max() + 1
    edges = np.linspace(0, t.max(), 21)
    for a, b in zip(edges[:-1], edges[1:]):
        m = (t >= a) & (t < b)
        per = [(int(A[m & (g == q)].mean()) if (m & (g == q)).any() else 0) for q in range(G)]
        print(f"  {a / 1e3:6.2f}-{b / 1e3:6.2f}s iters {m.sum():5d} active/group {per} TF/s {fl[m].sum() / ((b - a) * 1e-3) / 1e12:6.2f}")
