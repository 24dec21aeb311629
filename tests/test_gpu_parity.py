"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same inputs.

Tolerances (BASELINE.json north_star): neighbour index sets bit-exact; at fixed hyperparameters
posterior mean, std and NLML (and the gradient) within 1e-9 relative; with fitted hyperparameters
per-cell NLML within 1e-6 relative of the oracle's (or lower) and |d fs| <= 1 mm for most cells.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RTOL_FIXED = 1e-9


@pytest.fixture(scope="module")
def handle(small_day):
    import optimalinterpolation_b200 as oi
    h = oi.Handle(0)
    d = small_day
    h.set_observations(d.x_train, d.y_train, d.t_train, d.z)
    h.set_cells(d.X)
    h.gather_neighbours(d.radius_km * 1000.0)
    yield h
    h.close()


def test_neighbour_sets_bit_exact(handle, small_day, small_oracle):
    offsets, indices = handle.get_neighbours()
    o = small_oracle
    lists = o.tree.query_ball_point(small_day.X, r=o.radius_m)
    for c in range(len(small_day.X)):
        got = indices[offsets[c]:offsets[c + 1]]
        assert np.array_equal(got, np.sort(np.asarray(lists[c], dtype=np.int64))), c


def test_neighbour_lattice_ties():
    """Full 25 km lattice, centre on a lattice point, r = 300 km: 441 points incl. 4 exact ties
     -- the boundary is inclusive."""
    import optimalinterpolation_b200 as oi
    jj, ii = np.meshgrid(np.arange(40), np.arange(40))
    x = 25000.0 * jj.ravel(); y = 25000.0 * ii.ravel()
    h = oi.Handle(0)
    h.set_observations(x, y, np.zeros_like(x), np.zeros_like(x))
    h.set_cells(np.array([[25000.0 * 20, 25000.0 * 20], [25000.0 * 20 + 1.0, 25000.0 * 20]]))
    counts = h.gather_neighbours(300000.0)
    assert counts[0] == 441
    cx = 25000.0 * 20 + 1.0           # shifted centre: three of the four ties drop out
    assert counts[1] == np.count_nonzero((x - cx) ** 2 + (y - 25000.0 * 20) ** 2 <= 300000.0 ** 2) == 438
    h.close()


def test_neighbour_random_boundary(small_day):
    """Randomised near-boundary set equality against scipy's cKDTree (off-lattice coordinates)."""
    import optimalinterpolation_b200 as oi
    from scipy.spatial import cKDTree
    rng = np.random.default_rng(3)
    x = rng.uniform(0, 1e6, 5000); y = rng.uniform(0, 1e6, 5000)
    X = rng.uniform(2e5, 8e5, (300, 2))
    # put points (almost) exactly on the circle of the first cells
    th = rng.uniform(0, 2 * np.pi, 200)
    x[:200] = X[0, 0] + 150000.0 * np.cos(th); y[:200] = X[0, 1] + 150000.0 * np.sin(th)
    h = oi.Handle(0)
    h.set_observations(x, y, np.zeros_like(x), np.zeros_like(x)); h.set_cells(X)
    h.gather_neighbours(150000.0)
    offsets, indices = h.get_neighbours()
    lists = cKDTree(np.c_[x, y]).query_ball_point(X, r=150000.0)
    for c in range(len(X)):
        assert np.array_equal(indices[offsets[c]:offsets[c + 1]], np.sort(np.asarray(lists[c], dtype=np.int64)))
    h.close()


HYPER_SETS = {
    "x0": lambda d: d.x0,
    "notebook_optimum": lambda d: np.log([2.15e5, 1.40e5, 21.0, 0.0279, 0.00346, 0.1]),
    "flat_large_ell": lambda d: np.log([2.0e6, 3.0e6, 60.0, 0.5, 0.02, 0.1]),
    "short_ell": lambda d: np.log([4.0e4, 3.0e4, 2.0, 0.01, 0.002, 0.1]),
}


@pytest.mark.parametrize("name", list(HYPER_SETS))
def test_nlml_grad_fixed_hypers(handle, small_day, small_oracle, name):
    from oracle.gpr_oracle import nlml_grad
    d, o = small_day, small_oracle
    hyp = np.asarray(HYPER_SETS[name](d), dtype=float)
    nlz, grad = handle.nlml_grad(hyp, d.mean)
    cells = list(range(0, len(d.X), 7))
    worst_f = worst_g = 0.0
    for c in cells:
        _, inp, out, _ = o.cell_data(c, sort=True)
        f, g = nlml_grad(hyp, inp, out, np.ones(len(out)) * d.mean)
        worst_f = max(worst_f, abs(nlz[c] - f) / abs(f))
        gs = np.abs(g[:5]).max()
        worst_g = max(worst_g, np.abs(grad[c, :5] - g[:5]).max() / gs)
        assert grad[c, 5] == 0.0
    print(f"{name}: worst rel nlZ {worst_f:.2e}, worst rel grad {worst_g:.2e} over {len(cells)} cells")
    assert worst_f < RTOL_FIXED and worst_g < RTOL_FIXED


def test_nlml_grad_exact_convention(handle, small_day):
    hyp = np.log([2.15e5, 1.40e5, 21.0, 0.0279, 0.00346])
    f0, g0 = handle.nlml_grad(hyp, small_day.mean, grad_convention=0)
    f1, g1 = handle.nlml_grad(hyp, small_day.mean, grad_convention=1)
    assert np.array_equal(f0, f1)
    assert np.allclose(g1[:, 3:5] * 2, g0[:, 3:5], rtol=1e-15) and np.array_equal(g0[:, :3], g1[:, :3])
    # the EXACT convention is the true derivative: central finite differences on a few cells
    eps = 1e-5
    for k in range(5):
        hp, hm = hyp.copy(), hyp.copy(); hp[k] += eps; hm[k] -= eps
        fp, _ = handle.nlml_grad(hp, small_day.mean); fm, _ = handle.nlml_grad(hm, small_day.mean)
        fd = (fp - fm) / (2 * eps)
        sel = slice(0, None, 50)
        assert np.allclose(fd[sel], g1[sel, k], rtol=2e-5, atol=1e-5), k


def test_predict_fixed_hypers(handle, small_day, small_oracle):
    d, o = small_day, small_oracle
    nc = len(d.X)
    rng = np.random.default_rng(1)
    hyp = np.tile([2.15e5, 1.40e5, 21.0, 0.0279, 0.00346], (nc, 1)) * rng.uniform(0.5, 2.0, (nc, 5))
    p = handle.make_params(d.radius_km * 1000, d.T_mid, d.mean, d.x0, mode=1)
    handle.run(p, hyp)
    res = handle.get_results()
    worst = np.zeros(3)
    for c in range(0, nc, 5):
        ref = o.gpr3d(c, hypers=hyp[c], sort=True)
        got = res["out"][c]
        for q in range(3):
            worst[q] = max(worst[q], abs(got[q] - ref[q]) / abs(ref[q]))
        assert np.array_equal(got[3:], hyp[c])
    print("predict worst rel (fs, sfs2, lZ):", worst)
    assert (worst < RTOL_FIXED).all()


def test_cholesky_failure_semantics(small_day):
    """Duplicated points with vanishing noise: K is singular -> NaN tuple for that cell, inf NLML."""
    import optimalinterpolation_b200 as oi
    x = np.array([0.0, 0.0, 25000.0, 25000.0, 50000.0]); y = np.zeros(5); t = np.array([1.0, 1.0, 2.0, 2.0, 3.0])
    z = np.array([0.1, 0.1, 0.2, 0.2, 0.3])
    h = oi.Handle(0)
    h.set_observations(x, y, t, z); h.set_cells(np.array([[25000.0, 0.0], [9e6, 9e6]]))
    counts = h.gather_neighbours(1e5)
    assert list(counts) == [5, 0]
    hyp = np.log([25000.0, 25000.0, 1.0, 1.0, 1e-300])
    f, g = h.nlml_grad(hyp, 0.1)
    assert np.isinf(f[0]) and np.isinf(g[0]).all() and np.isnan(f[1])
    p = h.make_params(1e5, 2.0, 0.1, None, mode=1)
    h.run(p, np.tile(np.exp(hyp), (2, 1)))
    res = h.get_results()
    assert np.isnan(res["out"]).all() and list(res["status"]) == [3, 4]
    h.close()


def test_fit_matches_oracle(small_day):
    """Fitted path on the small day: device lockstep scipy-CG vs the reference path (tests/golden/small_day_fit.npz:
    scipy.optimize.minimize(CG) on the oracle objective for every 4th cell, in the reference's neighbour order and in
    ascending order).  Small cells (n ~ 40...300) are where the reference's line searches are most chaotic:
    the reference misses BASELINE.json's gate against ITSELF (tree order vs sorted order) on several per cent of them, so
    the gate is the reference's own miss rate + 0.1 % + 2 sigma of the sample; see tests/test_gpu_fit_parity.py for the
    full-size fixture."""
    import os
    import optimalinterpolation_b200 as oi
    from test_gpu_fit_parity import compare, summary
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "small_day_fit.npz")
    if not os.path.exists(path):
        pytest.skip("tests/golden/small_day_fit.npz not generated")
    fx = np.load(path)
    d = small_day
    cells = fx["cells"]
    g = oi.GPRDay(d.x_train, d.y_train, d.t_train, d.z, d.X[cells], d.radius_km, d.mean, d.T_mid, d.x0)
    res = g.run(opt=True)
    g.handle.close()
    floor = compare(fx["out_sorted"], fx["out_tree"])
    gpu = compare(res["out"], fx["out_tree"])
    N = len(cells)
    miss_floor, miss_gpu = 1 - floor["ok"].mean(), 1 - gpu["ok"].mean()
    sigma = np.sqrt(max(miss_floor, 1.0 / N) * (1 - miss_floor) / N)
    print(f"small-day fit parity over {N} cells: reference sorted-vs-tree {summary(floor)}\n GPU vs tree {summary(gpu)}\n"
          f" miss rate GPU {miss_gpu:.4f} reference floor {miss_floor:.4f} sigma {sigma:.4f}; nfev mean GPU {res['nfev'].mean():.1f} "
          f"ref {fx['nfev_tree'].mean():.1f}")
    import json
    out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    json.dump(dict(cells=int(N), reference_sorted_vs_tree=summary(floor), gpu_vs_tree=summary(gpu)),
              open(os.path.join(out_dir, "parity_small_day.json"), "w"), indent=1)
    assert miss_gpu <= miss_floor + 0.001 + 2 * sigma
    assert np.nanmedian(gpu["dfs"]) < 1e-3          # and the typical cell agrees to round-off


def test_groupings_bit_identical(small_day):
    """1 or 8 stream groups, tiny batches, forced express-lane hand-over, with and without CUDA-graph replay: fitted
    outputs, nfev and status must be bit-identical (=> results do not depend on batch composition, scheduling or GPU
    count).  (The experimental persistent engine and the flag-based fused Cholesky of round 1 were removed.)"""
    import optimalinterpolation_b200 as oi
    d = small_day
    cells = np.linspace(0, len(d.X) - 1, 30).round().astype(int)
    h = oi.Handle(0)
    h.set_observations(d.x_train, d.y_train, d.t_train, d.z); h.set_cells(d.X[cells]); h.gather_neighbours(d.radius_km * 1000.0)
    import os
    ref = None
    for kw in (dict(n_groups=1), dict(n_groups=8), dict(n_groups=3, max_active=7),
               dict(n_groups=4, express=(2, 5, 3)), dict(n_groups=2, nograph=True)):
        kw = dict(kw)
        ex = kw.pop("express", None)
        if kw.pop("nograph", False):     # the first configurations replay CUDA graphs (batches <= 32 cells); this one does not
            os.environ["OI_GRAPH_MAX"] = "0"
        if ex:   # force the express-lane hand-over (lanes, after-iterations, lane capacity) on this tiny problem
            os.environ.update(OI_EXPRESS=str(ex[0]), OI_EXPRESS_AFTER=str(ex[1]), OI_EXPRESS_CAP=str(ex[2]))
        try:
            h.run(h.make_params(d.radius_km * 1000.0, d.T_mid, d.mean, d.x0, mode=0, **kw))
        finally:
            for k in ("OI_EXPRESS", "OI_EXPRESS_AFTER", "OI_EXPRESS_CAP", "OI_GRAPH_MAX"):
                os.environ.pop(k, None)
        if ex:
            assert h.stats()["n_express_cells"] > 0
        if ref is None:
            assert h.stats()["n_graph_launches"] > 0
        r = h.get_results()
        if ref is None:
            ref = r
            assert np.isfinite(r["out"][:, 0]).mean() > 0.8
        else:
            assert np.array_equal(r["out"], ref["out"], equal_nan=True), kw
            assert np.array_equal(r["nfev"], ref["nfev"]) and np.array_equal(r["status"], ref["status"]), kw
    # the removed engine is refused, not silently replaced
    with pytest.raises(oi.gpr.OIError):
        h.run(h.make_params(d.radius_km * 1000.0, d.T_mid, d.mean, d.x0, mode=0, engine=1))
    h.close()


def test_ragged_block_edges_and_empty_cells():
    """Cells whose observation counts sit on the edges of the 64-row blocks and 16-wide DMMA chunks
    (n = 1, 2, 15..17, 63..65, 127..129, 191..193) plus cells with no observation at all: NLML, gradient and
    the posterior against the CPU oracle within 1e-9; empty cells give the NaN tuple and status NO_OBS."""
    import optimalinterpolation_b200 as oi
    from oracle.gpr_oracle import nlml_grad, predict
    rng = np.random.default_rng(11)
    sizes = [1, 2, 15, 16, 17, 63, 64, 65, 127, 128, 129, 191, 192, 193, 0, 0]
    # every cell gets its own cluster of observations, far (10 000 km apart) from every other cluster
    xs, ys, ts, zs, X = [], [], [], [], []
    for k, n in enumerate(sizes):
        cx, cy = 1.0e7 * k, -2.0e7 * (k % 3)
        X.append([cx, cy])
        r = 2.9e5 * np.sqrt(rng.uniform(0, 1, n)); th = rng.uniform(0, 2 * np.pi, n)
        xs.append(cx + r * np.cos(th)); ys.append(cy + r * np.sin(th))
        ts.append(rng.integers(0, 9, n).astype(float))
        zs.append(0.1 + 0.05 * np.sin(xs[-1] / 2e5) + 0.04 * rng.standard_normal(n))
    x, y, t, z = map(np.concatenate, (xs, ys, ts, zs))
    perm = rng.permutation(len(x)); x, y, t, z = x[perm], y[perm], t[perm], z[perm]
    X = np.array(X); mean = 0.1
    h = oi.Handle(0)
    h.set_observations(x, y, t, z); h.set_cells(X)
    counts = h.gather_neighbours(300000.0)
    assert list(counts) == sizes
    hyp = np.log([2.15e5, 1.40e5, 21.0, 0.0279, 0.00346, 0.1])
    nlz, grad = h.nlml_grad(hyp, mean)
    hn = np.tile(np.exp(hyp[:5]), (len(sizes), 1))
    h.run(h.make_params(300000.0, 4.0, mean, list(hyp), mode=1), hn)
    res = h.get_results()
    offsets, indices = h.get_neighbours()
    for k, n in enumerate(sizes):
        if n == 0:
            assert np.isnan(nlz[k]) and np.isnan(res["out"][k]).all() and res["status"][k] == 4
            continue
        ID = indices[offsets[k]:offsets[k + 1]]
        inp = np.c_[x[ID], y[ID], t[ID]]
        f, g = nlml_grad(hyp, inp, z[ID], np.ones(n) * mean)
        assert abs(nlz[k] - f) <= 1e-9 * max(abs(f), 1.0), (n, nlz[k], f)
        assert np.abs(grad[k] - g).max() <= 1e-9 * max(np.abs(g).max(), 1.0), (n, grad[k], g)
        fs, sfs2, lZ = predict(inp, z[ID], mean, np.array([[X[k, 0], X[k, 1], 4.0]]), list(hn[k, :3]), hn[k, 3], hn[k, 4])
        got = res["out"][k]
        assert abs(got[0] - fs) <= 1e-9 * abs(fs) and abs(got[1] - sfs2) <= 1e-9 * abs(sfs2) and abs(got[2] - lZ) <= 1e-9 * max(abs(lZ), 1.0), (n, got[:3], (fs, sfs2, lZ))
    h.close()


def test_first_cg_iterations_match_scipy(small_day, small_oracle):
    """Before round-off differences can steer a line search elsewhere, the device optimiser must walk scipy's path:
    with maxiter = 3 (scipy options={'maxiter': 3}) the number of evaluations and the warn flag are equal and the
    hyperparameters agree to 1e-7; same with a loose gtol that stops some cells early with status 0."""
    import warnings
    import scipy.optimize
    import optimalinterpolation_b200 as oi
    from oracle.gpr_oracle import nlml_grad
    d, o = small_day, small_oracle
    cells = np.arange(3, len(d.X), 61)
    h = oi.Handle(0)
    h.set_observations(d.x_train, d.y_train, d.t_train, d.z); h.set_cells(d.X[cells]); h.gather_neighbours(d.radius_km * 1000.0)
    for opts, kw in (({"maxiter": 3}, dict(maxiter=3)), ({"maxiter": 6, "gtol": 5.0}, dict(maxiter=6, gtol=5.0))):
        h.run(h.make_params(d.radius_km * 1000.0, d.T_mid, d.mean, d.x0, mode=0, **kw))
        res = h.get_results()
        statuses = set()
        for k, c in enumerate(cells):
            _, inp, out, _ = o.cell_data(int(c), sort=True)
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                r = scipy.optimize.minimize(nlml_grad, list(d.x0), args=(inp, out, np.ones(len(out)) * d.mean), method="CG",
                                            jac=True, options=opts)
            assert res["nfev"][k] == r.nfev and res["status"][k] == r.status, (c, opts, res["nfev"][k], r.nfev, res["status"][k], r.status)
            assert np.allclose(res["out"][k, 3:8], np.exp(r.x[:5]), rtol=1e-7, atol=0), (c, opts)
            statuses.add(int(r.status))
        print(opts, "statuses seen", statuses, "nfev", res["nfev"].tolist())
    h.close()
