"""GPU: FITTED parity at BASELINE.json's gate -- per-cell NLML at the optimum within 1e-6 relative of the reference's (or
lower) and fitted freeboard within 1 mm for >= 99.9 % of cells -- on fixtures of the reference path (oracle = restatement
of GPR3D, bit-identical to the reference functions, tests/test_oracle.py) that cover the WHOLE n range of the day:

  tests/golden/day_fit_sample_1k.npz   1024 cells at the 1024-quantiles of the day's n distribution (158...1786), each fitted
                                       in the reference's neighbour order (``*_tree``) and in ascending index order
                                       (``*_sorted``, the order the CUDA path uses)
  tests/golden/cfg5_fit_sample.npz     cells of the 12.5 km / 500 km workload (BASELINE.json configs[4], n 1500...4900)

The reference's stopping point is not invariant to a permutation of its own inputs (the line searches
amplify round-off), so the reference misses its own gate against itself; that miss rate, measured on the same cells
(tree order vs sorted order), is the floor, and the gate is "GPU miss rate <= floor + 0.1 % (+ 2 sigma of the sample)".
Every number is printed and written to gpurun_out/parity_*.json.
"""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
FIX_1K = os.path.join(HERE, "golden", "day_fit_sample_1k.npz")
FIX_CFG5 = os.path.join(HERE, "golden", "cfg5_fit_sample.npz")
OUT_DIR = os.path.join(os.path.dirname(HERE), "gpurun_out")


def compare(a, b):
    """a (candidate) against b (reference), rows = the reference's 8-tuple (fs, std, lZ, hypers).  Returns per-cell flags."""
    nan_a, nan_b = np.isnan(a[:, 0]), np.isnan(b[:, 0])
    both = ~nan_a & ~nan_b
    dfs = np.full(len(a), np.nan); rel = np.full(len(a), np.nan)
    dfs[both] = np.abs(a[both, 0] - b[both, 0]) * 1e3                       # mm
    rel[both] = (a[both, 2] - b[both, 2]) / np.abs(b[both, 2])             # lZ = -NLML: >= -1e-6 means "within 1e-6 or lower NLML"
    ok_fs = both & (dfs <= 1.0)
    ok_nl = both & (rel >= -1e-6)
    ok = (ok_fs & ok_nl) | (nan_a & nan_b)
    return dict(nan_a=nan_a, nan_b=nan_b, both=both, dfs=dfs, rel=rel, ok_fs=ok_fs | (nan_a & nan_b), ok_nl=ok_nl | (nan_a & nan_b), ok=ok)


def summary(c):
    n = len(c["ok"]); b = c["both"]
    return dict(cells=int(n), finite_both=int(b.sum()), nan_only_candidate=int((c["nan_a"] & ~c["nan_b"]).sum()),
                nan_only_reference=int((~c["nan_a"] & c["nan_b"]).sum()), nan_both=int((c["nan_a"] & c["nan_b"]).sum()),
                frac_fs_1mm=float(c["ok_fs"].mean()), frac_nlml=float(c["ok_nl"].mean()), frac_gate=float(c["ok"].mean()),
                dfs_mm_median=float(np.nanmedian(c["dfs"])), dfs_mm_p99=float(np.nanpercentile(c["dfs"], 99)), dfs_mm_max=float(np.nanmax(c["dfs"])),
                frac_fs_1mm_of_finite=float((c["dfs"][b] <= 1.0).mean()), frac_nlml_of_finite=float((c["rel"][b] >= -1e-6).mean()))


@pytest.fixture(scope="module")
def day():
    from optimalinterpolation_b200.synthetic import make_day
    return make_day()


@pytest.fixture(scope="module")
def fixture_1k():
    if not os.path.exists(FIX_1K):
        pytest.skip("tests/golden/day_fit_sample_1k.npz not generated")
    f = np.load(FIX_1K)
    use = f["done_tree"] & f["done_sorted"]
    if use.sum() < 64:
        pytest.skip("fixture holds fewer than 64 finished cells")
    return {k: (f[k][use] if f[k].shape[:1] == use.shape else f[k]) for k in f.files}


@pytest.fixture(scope="module")
def gpu_fit_1k(day, fixture_1k):
    import optimalinterpolation_b200 as oi
    gd = oi.GPRDay(day.x_train, day.y_train, day.t_train, day.z, day.X[fixture_1k["cells"]], day.radius_km, day.mean, day.T_mid, day.x0)
    res = gd.run(opt=True)
    st = gd.handle.stats()
    gd.handle.close()
    return res, st


def test_fit_parity_whole_n_range(day, fixture_1k, gpu_fit_1k):
    fx = fixture_1k
    res, _ = gpu_fit_1k
    assert np.array_equal(res["n"], fx["n"])
    N = len(fx["cells"])
    floor = compare(fx["out_sorted"], fx["out_tree"])          # the reference against itself (permuted inputs)
    g_tree = compare(res["out"], fx["out_tree"])               # CUDA path against the reference as the reference runs it
    g_sort = compare(res["out"], fx["out_sorted"])             # ... against the reference fed the CUDA path's input order
    rep = dict(cells=N, n_min=int(fx["n"].min()), n_max=int(fx["n"].max()), frac_cells_n_gt_1100=float((fx["n"] > 1100).mean()),
               reference_sorted_vs_tree=summary(floor), gpu_vs_tree=summary(g_tree), gpu_vs_sorted=summary(g_sort),
               nfev_mean=dict(gpu=float(res["nfev"].mean()), ref_tree=float(fx["nfev_tree"].mean()), ref_sorted=float(fx["nfev_sorted"].mean())),
               nfev_identical_to_sorted_run=float((res["nfev"] == fx["nfev_sorted"]).mean()),
               out_bitwise_identical_to_sorted_run=float(np.all((res["out"] == fx["out_sorted"]) | (np.isnan(res["out"]) & np.isnan(fx["out_sorted"])), axis=1).mean()))
    # non-finite outcomes by n (the day product's holes): both sides, per quartile of n
    edges = np.quantile(fx["n"], [0, .25, .5, .75, 1.0])
    rep["nonfinite_by_n_quartile"] = []
    for q in range(4):
        m = (fx["n"] >= edges[q]) & ((fx["n"] <= edges[q + 1]) if q == 3 else (fx["n"] < edges[q + 1]))
        rep["nonfinite_by_n_quartile"].append(dict(
            n_lo=int(edges[q]), n_hi=int(edges[q + 1]), cells=int(m.sum()), gpu=float(np.isnan(res["out"][m, 0]).mean()),
            ref_tree=float(np.isnan(fx["out_tree"][m, 0]).mean()), ref_sorted=float(np.isnan(fx["out_sorted"][m, 0]).mean()),
            gpu_status_hist=np.bincount(res["status"][m], minlength=6).tolist(),
            ref_tree_status_hist=np.bincount(np.clip(fx["status_tree"][m], 0, 5), minlength=6).tolist()))
    os.makedirs(OUT_DIR, exist_ok=True)
    json.dump(rep, open(os.path.join(OUT_DIR, "parity_1k.json"), "w"), indent=1)
    print(json.dumps(rep, indent=1))
    miss_floor, miss_gpu = 1 - floor["ok"].mean(), 1 - g_tree["ok"].mean()
    sigma = np.sqrt(max(miss_floor, 1.0 / N) * (1 - miss_floor) / N)
    assert miss_gpu <= miss_floor + 0.001 + 2 * sigma, (miss_gpu, miss_floor, sigma)
    # the holes of the day product: the CUDA path must not lose more cells than the reference loses against itself
    nan_gpu, nan_ref, nan_srt = np.isnan(res["out"][:, 0]).mean(), np.isnan(fx["out_tree"][:, 0]).mean(), np.isnan(fx["out_sorted"][:, 0]).mean()
    assert nan_gpu <= max(nan_ref, nan_srt) + 0.001 + 2 * np.sqrt(max(nan_ref, 1.0 / N) / N), (nan_gpu, nan_ref, nan_srt)
    # against the reference run in the SAME input order, most cells agree to round-off
    assert np.nanmedian(g_sort["dfs"]) < 1e-3


def test_lbfgs_fast_mode_reaches_reference_nlml(day, fixture_1k, gpu_fit_1k):
    """The fast mode (exact gradient + L-BFGS, oi_params.optimiser = OI_OPT_LBFGS) is NOT the parity mode: it is judged by
    the likelihood it reaches (>= the reference's, i.e. NLML <= reference * (1 + 1e-6)) and by its evaluation count."""
    import optimalinterpolation_b200 as oi
    fx = fixture_1k
    res_cg, st_cg = gpu_fit_1k
    gd = oi.GPRDay(day.x_train, day.y_train, day.t_train, day.z, day.X[fx["cells"]], day.radius_km, day.mean, day.T_mid, day.x0,
                   grad_convention=1)
    res = gd.run(opt=True, optimiser=1)
    st = gd.handle.stats()
    gd.handle.close()
    ref = fx["out_tree"]
    fin_ref = np.isfinite(ref[:, 0]); fin = np.isfinite(res["out"][:, 0])
    m = fin & fin_ref
    rel = (res["out"][m, 2] - ref[m, 2]) / np.abs(ref[m, 2])
    dfs = np.abs(res["out"][m, 0] - ref[m, 0]) * 1e3
    rep = dict(cells=int(len(fin)), finite_fast=float(fin.mean()), finite_ref=float(fin_ref.mean()), finite_cg_gpu=float(np.isfinite(res_cg["out"][:, 0]).mean()),
               frac_lZ_ge_ref=float((rel >= -1e-6).mean()), frac_lZ_strictly_higher_1e_6=float((rel > 1e-6).mean()),
               dfs_mm_median=float(np.median(dfs)), dfs_mm_p90=float(np.percentile(dfs, 90)), dfs_mm_p99=float(np.percentile(dfs, 99)),
               frac_fs_1mm=float((dfs <= 1.0).mean()), frac_fs_5mm=float((dfs <= 5.0).mean()),
               nfev_mean_fast=float(res["nfev"].mean()), nfev_max_fast=int(res["nfev"].max()), nfev_mean_ref=float(fx["nfev_tree"].mean()),
               status_hist_fast=np.bincount(res["status"], minlength=6).tolist(),
               device_ms_fast=float(st["ms_total"]), device_ms_cg=float(st_cg["ms_total"]),
               cells_per_s_fast=float(len(fin) / st["ms_total"] * 1e3), cells_per_s_cg=float(len(fin) / st_cg["ms_total"] * 1e3))
    os.makedirs(OUT_DIR, exist_ok=True)
    json.dump(rep, open(os.path.join(OUT_DIR, "parity_lbfgs_1k.json"), "w"), indent=1)
    print(json.dumps(rep, indent=1))
    assert fin.mean() >= fin_ref.mean()                      # no more holes than the reference
    assert (rel >= -1e-6).mean() >= 0.98
    assert res["nfev"].mean() < 0.5 * fx["nfev_tree"].mean()


def test_cfg5_fit_parity():
    """BASELINE.json configs[4] (12.5 km lattice, 500 km radius): fits at n = 1500...4900 against the reference path."""
    if not os.path.exists(FIX_CFG5):
        pytest.skip("tests/golden/cfg5_fit_sample.npz not generated")
    import optimalinterpolation_b200 as oi
    from optimalinterpolation_b200.synthetic import make_day_cfg5
    f = np.load(FIX_CFG5)
    use = f["done"]
    if use.sum() < 4:
        pytest.skip("fewer than 4 finished cells")
    d = make_day_cfg5()
    cells, ref = f["cells"][use], f["out"][use]
    gd = oi.GPRDay(d.x_train, d.y_train, d.t_train, d.z, d.X[cells], d.radius_km, d.mean, d.T_mid, d.x0)
    res = gd.run(opt=True)
    st = gd.handle.stats()
    gd.handle.close()
    assert np.array_equal(res["n"], f["n"][use])
    c = compare(res["out"], ref)
    rep = dict(summary(c), n=f["n"][use].tolist(), dfs_mm=[None if np.isnan(v) else float(v) for v in c["dfs"]],
               rel_lZ=[None if np.isnan(v) else float(v) for v in c["rel"]], nfev_gpu=res["nfev"].tolist(), nfev_ref=f["nfev"][use].tolist(),
               status_gpu=res["status"].tolist(), status_ref=f["status"][use].tolist(), device_s=float(st["ms_total"] * 1e-3),
               tflops=float(st["flops"] / st["ms_total"] * 1e-9))
    os.makedirs(OUT_DIR, exist_ok=True)
    json.dump(rep, open(os.path.join(OUT_DIR, "parity_cfg5.json"), "w"), indent=1)
    print(json.dumps(rep, indent=1))
    # one cell of the sample may land elsewhere; the others must meet BASELINE.json's tolerances
    assert c["ok"].sum() >= len(cells) - 1


def test_two_handles_in_one_process(small_day):
    """oi_create sets the > 48 KB shared-memory opt-in per device: a second handle (on a second GPU when the box has one,
    else on the same GPU) must fit too, and both must give the same numbers."""
    import optimalinterpolation_b200 as oi
    d = small_day
    cells = np.arange(0, len(d.X), 60)
    h0 = oi.Handle(0)
    try:
        h1 = oi.Handle(1)
        second = "cuda:1"
    except oi.gpr.OIError:
        h1 = oi.Handle(0)
        second = "cuda:0 (single-GPU box)"
    outs = []
    for h in (h0, h1):
        h.set_observations(d.x_train, d.y_train, d.t_train, d.z); h.set_cells(d.X[cells]); h.gather_neighbours(d.radius_km * 1000.0)
        h.run(h.make_params(d.radius_km * 1000.0, d.T_mid, d.mean, d.x0, mode=0))
        outs.append(h.get_results()["out"])
    h0.close(); h1.close()
    print("second handle on", second)
    assert np.array_equal(outs[0], outs[1], equal_nan=True) and np.isfinite(outs[0][:, 0]).mean() > 0.8


def test_shared_work_list_two_handles(small_day):
    """Two handles (two host threads, as two GPU processes would) draw from ONE shared cost-sorted work list
    (oi_set_shared_queue): every cell is computed by exactly one of them and the merged field is bit-identical to a
    single handle's (a cell's numbers do not depend on who computes it)."""
    import threading
    import optimalinterpolation_b200 as oi
    d = small_day
    cells = np.arange(0, len(d.X), 3)
    X = d.X[cells]
    ref_h = oi.Handle(0)
    # few slots per handle: admission is greedy up to a handle's capacity, and on this small problem an unlimited handle
    # would claim the whole list before the other thread has started
    p = ref_h.make_params(d.radius_km * 1000.0, d.T_mid, d.mean, d.x0, mode=0, max_active=12)
    ref = ref_h.gpr_day(d.x_train, d.y_train, d.t_train, d.z, X, p)
    assert ref_h.get_owned().all()
    ref_h.close()
    name = f"/oi_b200_gputest_{os.getpid()}"
    hs = [oi.Handle(0), oi.Handle(0)]
    for h in hs:
        h.set_shared_queue(name)
    outs = [None, None]

    def work(k):
        for _ in range(2):                                   # two consecutive runs: the list's generation advances on both
            r = hs[k].gpr_day(d.x_train, d.y_train, d.t_train, d.z, X, p)
        outs[k] = (r, hs[k].get_owned())
    th = [threading.Thread(target=work, args=(k,)) for k in range(2)]
    [t.start() for t in th]; [t.join() for t in th]
    (r0, o0), (r1, o1) = outs
    empty = r0["n"] == 0
    assert np.array_equal(o0 & o1, empty)                    # only cells without observations are "owned" by both
    print("cells owned by the two handles:", int(o0.sum()), int(o1.sum()), "of", len(cells))
    assert (o0 | o1).all()          # (how the cells split between the two depends on thread timing: not asserted)
    merged = np.where(o0[:, None], r0["out"], r1["out"])
    assert np.array_equal(merged, ref["out"], equal_nan=True)
    assert np.array_equal(np.where(o0, r0["nfev"], r1["nfev"]), ref["nfev"])
    hs[0].set_shared_queue(None); hs[1].set_shared_queue(None)
    hs[0].unlink_shared_queue(name)
    [h.close() for h in hs]
