"""Scheduler experiments on the bench step (2 stripes = 2390 cells of the day): one fit per setting of the OI_* environment
switches that run_lockstep reads, device ms and the evaluation/iteration counts.  (needs a GPU)
usage: python tools/sched_experiment.py "NAME=VAL,NAME=VAL" "NAME=VAL" ...   ('' = defaults)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import numpy as np
import optimalinterpolation_b200 as oi
from optimalinterpolation_b200.synthetic import make_day
d = make_day()
nc = len(d.X)
cells = np.sort(np.concatenate([np.arange(s, nc, 16) for s in range(2)]))
h = oi.Handle(0)
h.set_observations(d.x_train, d.y_train, d.t_train, d.z); h.set_cells(d.X[cells]); h.gather_neighbours(d.radius_km * 1000.0)
p = h.make_params(d.radius_km * 1000.0, d.T_mid, d.mean, d.x0, mode=0)
ref = None
for k, spec in enumerate([""] + sys.argv[1:]):
    env = dict(kv.split("=") for kv in spec.split(",") if kv)
    os.environ.update(env)
    t0 = time.perf_counter(); h.run(p); wall = time.perf_counter() - t0
    for name in env:
        os.environ.pop(name, None)
    st = h.stats(); r = h.get_results()
    if ref is None:
        ref = r
    same = np.array_equal(r["out"], ref["out"], equal_nan=True)
    print(f"{'warm-up (defaults)' if k == 0 else (spec or 'defaults'):60s} device ms {st['ms_total']:9.1f} wall {wall:7.2f} s iterations {st['n_iterations']:6d} "
          f"graph launches {st['n_graph_launches']:6d} express cells {st['n_express_cells']:4d} identical {same}", flush=True)
