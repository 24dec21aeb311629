"""GPU tests at BASELINE.json's full sizes (cells of the synthetic 25 km pan-Arctic day, n up to ~1800) and of the
two-pass day product.  Size-independent properties where the CPU oracle would take too long.  (Fitted parity at full size:
tests/test_gpu_fit_parity.py.)"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def day():
    from optimalinterpolation_b200.synthetic import make_day
    return make_day()


@pytest.fixture(scope="module")
def big_handle(day):
    import optimalinterpolation_b200 as oi
    h = oi.Handle(0)
    h.set_observations(day.x_train, day.y_train, day.t_train, day.z)
    yield h
    h.close()


HYP = np.log([2.15e5, 1.40e5, 21.0, 0.0279, 0.00346, 0.1])


def test_full_size_smlii_vs_oracle(day, big_handle):
    """n = min / median / max of the day: NLML + gradient within 1e-9 of the CPU oracle."""
    from oracle.gpr_oracle import DayOracle, nlml_grad
    h = big_handle
    h.set_cells(day.X)
    counts = h.gather_neighbours(day.radius_km * 1000.0)
    order = np.argsort(counts)
    cells = np.array([order[0], order[len(order) // 2], order[-1], order[len(order) // 4]])
    h.set_cells(day.X[cells]); h.gather_neighbours(day.radius_km * 1000.0)
    nlz, grad = h.nlml_grad(HYP, day.mean)
    o = DayOracle.from_day(day)
    for k, c in enumerate(cells):
        _, inp, out, _ = o.cell_data(int(c), sort=True)
        f, g = nlml_grad(HYP, inp, out, np.ones(len(out)) * day.mean)
        assert abs(nlz[k] - f) <= 1e-9 * abs(f), (len(out), nlz[k], f)
        assert np.abs(grad[k] - g).max() <= 1e-9 * np.abs(g).max(), (len(out), grad[k], g)
    print("n of the checked cells:", counts[cells])


def test_batch_composition_invariance(day, big_handle):
    """A cell's numbers must not depend on which other cells share the launch (=> GPU count independent)."""
    h = big_handle
    rng = np.random.default_rng(5)
    big = rng.choice(len(day.X), 300, replace=False)
    sub = big[::8]
    h.set_cells(day.X[big]); h.gather_neighbours(day.radius_km * 1000.0)
    f_big, g_big = h.nlml_grad(HYP, day.mean)
    h.set_cells(day.X[sub][::-1].copy()); h.gather_neighbours(day.radius_km * 1000.0)
    f_sub, g_sub = h.nlml_grad(HYP, day.mean)
    assert np.array_equal(f_big[::8], f_sub[::-1]) and np.array_equal(g_big[::8], g_sub[::-1])


def test_shift_invariance_and_finite_differences(day, big_handle):
    """z -> z + c with mean -> mean + c leaves NLML/gradient unchanged and shifts the posterior mean by c;
    the EXACT gradient convention matches central finite differences at full size."""
    import optimalinterpolation_b200 as oi
    h = big_handle
    cells = np.arange(0, len(day.X), 997)
    h.set_cells(day.X[cells]); h.gather_neighbours(day.radius_km * 1000.0)
    f0, g0 = h.nlml_grad(HYP, day.mean, grad_convention=1)
    eps = 1e-5
    for k in range(5):
        hp, hm = HYP.copy(), HYP.copy(); hp[k] += eps; hm[k] -= eps
        fp, _ = h.nlml_grad(hp, day.mean); fm, _ = h.nlml_grad(hm, day.mean)
        assert np.allclose((fp - fm) / (2 * eps), g0[:, k], rtol=5e-5, atol=1e-4), k
    hyp_nat = np.tile(np.exp(HYP[:5]), (len(cells), 1))
    p = h.make_params(day.radius_km * 1000, day.T_mid, day.mean, day.x0, mode=1)
    h.run(p, hyp_nat); r0 = h.get_results()["out"].copy()
    c = 0.25                                        # exactly representable shift
    h2 = oi.Handle(0)
    h2.set_observations(day.x_train, day.y_train, day.t_train, day.z + c)
    h2.set_cells(day.X[cells]); h2.gather_neighbours(day.radius_km * 1000.0)
    f1, g1 = h2.nlml_grad(HYP, day.mean + c, grad_convention=1)
    assert np.allclose(f1, f0, rtol=1e-9) and np.allclose(g1, g0, rtol=1e-7, atol=1e-7)
    p2 = h2.make_params(day.radius_km * 1000, day.T_mid, day.mean + c, day.x0, mode=1)
    h2.run(p2, hyp_nat); r1 = h2.get_results()["out"]
    assert np.allclose(r1[:, 0] - c, r0[:, 0], rtol=0, atol=1e-9) and np.allclose(r1[:, 1:3], r0[:, 1:3], rtol=1e-8)
    h2.close()


def test_two_pass_small_day(small_day, small_oracle):
    """Pass 1 -> smoothing -> pass 2 (GPR_CS2S3.py:299-334): pass-2 predictions equal the oracle's GPR3D(opt=False)
    at the same smoothed hyperparameters."""
    import optimalinterpolation_b200 as oi
    from optimalinterpolation_b200.postprocess import two_pass
    d, o = small_day, small_oracle
    gd = oi.GPRDay.from_day(d)
    sie = np.full(d.shape, np.nan); sie[d.ids] = 1.0
    res = two_pass(gd, d.ids, d.shape, sie, date="20190128", grid_res=d.grid_res_km, T=d.T)
    keys = {"20190128" + s for s in ("_interp", "_interp_error", "_lZ", "_ell_x", "_ell_y", "_ell_t", "_sf2", "_sn2",
                                      "_ell_x_smth", "_ell_y_smth", "_ell_t_smth", "_sf2_smth", "_sn2_smth",
                                      "_interp_smth", "_interp_error_smth")}
    assert keys <= set(res)
    ell = np.array([res["20190128_ell_x_smth"][d.ids], res["20190128_ell_y_smth"][d.ids], res["20190128_ell_t_smth"][d.ids]]).T
    sf2, sn2 = res["20190128_sf2_smth"][d.ids], res["20190128_sn2_smth"][d.ids]
    assert np.isfinite(ell).all() and (ell[:, 0] <= 2 * d.radius_km * 1000 + 1e-6).all() and (sf2 <= 0.1 + 1e-12).all()
    fs, er = res["20190128_interp_smth"][d.ids], res["20190128_interp_error_smth"][d.ids]
    for c in range(0, len(d.X), 25):
        ref = o.gpr3d(c, hypers=[ell[c, 0], ell[c, 1], ell[c, 2], sf2[c], sn2[c]], sort=True)
        assert abs(fs[c] - ref[0]) <= 1e-9 * abs(ref[0]) and abs(er[c] - ref[1]) <= 1e-8 * abs(ref[1]), c


def test_config5_large_cell_vs_oracle():
    """BASELINE.json configs[4]: 12.5 km lattice, r = 500 km, thousands of observations per cell (here one cell with
    n ~ 3300, N = 52 blocks): NLML, gradient and posterior within 1e-9 of the CPU oracle; the neighbour set is exact."""
    import optimalinterpolation_b200 as oi
    from oracle.gpr_oracle import nlml_grad, predict, neighbours_brute
    rng = np.random.default_rng(12)
    side = 100                                           # 12.5 km lattice sites around the cell, ~8 % of site-days observed
    jj, ii = np.meshgrid(np.arange(-side, side + 1), np.arange(-side, side + 1))
    xs, ys, ts = [], [], []
    for day in range(9):
        hit = rng.uniform(size=jj.shape) < 0.073
        xs.append(12500.0 * jj[hit]); ys.append(12500.0 * ii[hit]); ts.append(np.full(hit.sum(), float(day)))
    x, y, t = map(np.concatenate, (xs, ys, ts))
    z = 0.1 + 0.06 * np.sin(x / 3e5) * np.cos(y / 2.5e5) + 0.003 * (t - 4) + 0.04 * rng.standard_normal(x.size)
    X = np.array([[0.0, 0.0]]); mean = float(np.round(z.mean(), 3))
    h = oi.Handle(0)
    h.set_observations(x, y, t, z); h.set_cells(X)
    n = int(h.gather_neighbours(500000.0)[0])
    assert 2500 < n < 5000
    offsets, indices = h.get_neighbours()
    assert np.array_equal(indices, neighbours_brute(x, y, X[0], 500000.0))
    hyp = np.log([3.0e5, 2.2e5, 15.0, 0.02, 0.004, 0.1])
    nlz, grad = h.nlml_grad(hyp, mean)
    inp = np.c_[x[indices], y[indices], t[indices]]
    f, g = nlml_grad(hyp, inp, z[indices], np.ones(n) * mean)
    assert abs(nlz[0] - f) <= 1e-9 * abs(f), (n, nlz[0], f)
    assert np.abs(grad[0] - g).max() <= 1e-9 * np.abs(g).max(), (grad[0], g)
    hn = np.exp(hyp[:5])[None, :]
    h.run(h.make_params(500000.0, 4.0, mean, list(hyp), mode=1), hn)
    got = h.get_results()["out"][0]
    fs, sfs2, lZ = predict(inp, z[indices], mean, np.array([[0.0, 0.0, 4.0]]), list(hn[0, :3]), hn[0, 3], hn[0, 4])
    assert abs(got[0] - fs) <= 1e-9 * abs(fs) and abs(got[1] - sfs2) <= 1e-9 * abs(sfs2) and abs(got[2] - lZ) <= 1e-9 * abs(lZ)
    print("config-5 cell: n =", n, "rel err nlZ", abs(nlz[0] - f) / abs(f), "grad", np.abs(grad[0] - g).max() / np.abs(g).max())
    h.close()
