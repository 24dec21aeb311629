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
