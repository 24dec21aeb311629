"""CPU: the synthetic day follows the reference's flattening order and lattice."""
import numpy as np

from optimalinterpolation_b200.synthetic import make_day, make_small_day


def test_flatten_order_matches_reference_loop():
    d = make_small_day(seed=3)
    d2 = make_day(n_side=40, ice_radius_cells=14.0, centre=(20, 20), radius_km=100.0, tracks_per_day=4, seed=3,
                  s3_hole_cells=4.0, cs2_hole_cells=1.0, keep_sat=True)
    assert np.array_equal(d.z, d2.z)
    sat = d2.sat
    res = 25000.0
    jj, ii = np.meshgrid(np.arange(40), np.arange(40))
    x, y = res * jj, res * ii
    # the reference's loop, GPR_CS2S3.py:223-241
    xs = [[] for _ in range(4)]; ys = [[] for _ in range(4)]; ts = [[] for _ in range(4)]; zs = [[] for _ in range(4)]
    for day in range(sat.shape[3]):
        for s in range(4):
            ids = np.where(~np.isnan(sat[:, :, s, day]))
            xs[s].extend(x[ids]); ys[s].extend(y[ids]); ts[s].extend(np.ones(np.shape(ids)[1]) * day)
            zs[s].extend(sat[:, :, s, day][ids])
    assert np.array_equal(np.concatenate(xs), d2.x_train) and np.array_equal(np.concatenate(ys), d2.y_train)
    assert np.array_equal(np.concatenate(ts), d2.t_train) and np.array_equal(np.concatenate(zs), d2.z)


def test_full_day_shape():
    d = make_day()
    assert len(d.X) == 19109 and d.T == 9 and d.T_mid == 4 and len(d.x0) == 6
    assert np.all(d.x_train % 25000.0 == 0) and np.all(d.z >= -0.37) and np.all(d.z <= 0.63)
    assert d.z.size == 38144


def test_cfg5_day_geometry():
    """BASELINE.json configs[4]: 12.5 km lattice, 500 km radius -- thousands of observations per cell."""
    from scipy.spatial import cKDTree
    from optimalinterpolation_b200.synthetic import make_day_cfg5
    d = make_day_cfg5()
    assert d.shape == (640, 640) and d.grid_res_km == 12.5 and d.radius_km == 500.0 and len(d.X) == 76425
    assert np.allclose(d.x0[:2], np.log(12500.0))                      # x0 follows grid_res (GPR_CS2S3.py:217)
    cnt = np.asarray(cKDTree(np.c_[d.x_train, d.y_train]).query_ball_point(d.X[::256], r=500e3, return_length=True))
    assert 1000 < cnt.min() and 3000 < np.median(cnt) < 4200 and cnt.max() < 6000


def test_bench_config_is_one_object_for_both_arms():
    """bench.py builds `config` in one function from the workload alone, so the reference arm's line carries the GPU
    arm's config (the driver compares them) at every N."""
    import importlib.util, os, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(root, "bench.py"))
    b = importlib.util.module_from_spec(spec)
    argv, sys.argv = sys.argv, ["bench.py"]
    try:
        spec.loader.exec_module(b)
        args = b.parse()
    finally:
        sys.argv = argv
    for world in (1, 8):
        day, cells = b.make_workload(args, world)
        counts = np.arange(len(cells)) % 1500 + 200
        c1, parts = b.make_config(args, world, day, cells, counts)
        c2, _ = b.make_config(args, world, day, cells, counts)
        assert c1 == c2 and c1["cells_per_step"] == len(cells) and sum(len(p) for p in parts) == len(cells)
        assert ("dynamic" in c1["sharding"]) == (world > 1)
    assert len(b.make_workload(args, 8)[1]) == 19109                    # N = 8: one step is the whole day
