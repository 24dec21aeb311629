"""Season/day setup (GPR_CS2S3.py:201-246): CPU checks of the window flattening, GPU check of the season loop."""
import numpy as np
import pytest

from optimalinterpolation_b200.synthetic import make_day


def _small_season(T_total=11):
    d = make_day(n_side=40, ice_radius_cells=14.0, centre=(20, 20), radius_km=100.0, tracks_per_day=4, seed=5,
                 s3_hole_cells=4.0, cs2_hole_cells=1.0, keep_sat=True, T=T_total)
    res = 25000.0
    jj, ii = np.meshgrid(np.arange(40), np.arange(40))
    x, y = res * jj.astype(float), res * ii.astype(float)
    sie = np.full((40, 40, T_total), np.nan)
    sie[d.ids[0], d.ids[1], :] = 1.0
    sie[d.ids[0][::7], d.ids[1][::7], 6] = np.nan          # the ice mask changes from day to day
    return d.sat, sie, x, y


def test_flatten_window_matches_reference_loop():
    from optimalinterpolation_b200.season import flatten_window, day_inputs
    obs, sie, x, y = _small_season()
    for day in (0, 2):
        sat = obs[:, :, :, day:day + 9]                     # GPR_CS2S3.py:213
        # the reference's loop, GPR_CS2S3.py:223-241
        xs = [[] for _ in range(4)]; ys = [[] for _ in range(4)]; ts = [[] for _ in range(4)]; zs = [[] for _ in range(4)]
        for dd in range(sat.shape[3]):
            for s in range(4):
                ids = np.where(~np.isnan(sat[:, :, s, dd]))
                xs[s].extend(x[ids]); ys[s].extend(y[ids]); ts[s].extend(np.ones(np.shape(ids)[1]) * dd)
                zs[s].extend(sat[:, :, s, dd][ids])
        xt, yt, tt, z = flatten_window(obs, x, y, day, 9)
        assert np.array_equal(np.concatenate(xs), xt) and np.array_equal(np.concatenate(ys), yt)
        assert np.array_equal(np.concatenate(ts), tt) and np.array_equal(np.concatenate(zs), z)
        g = day_inputs(obs, sie, x, y, day, 9)
        ids = np.where(~np.isnan(sie[:, :, day + 4]))       # :214, :243-244
        assert np.array_equal(g["X"], np.array([x[ids], y[ids]]).T) and g["T_mid"] == 4
        assert g["mean"] == float(np.round(np.mean(z), 3))
    assert day_inputs(obs, sie, x, y, 0, 9, prior_mean=0.25)["mean"] == 0.25
    assert day_inputs(obs, sie, x, y, 2, 9, prior_mean=lambda d: 0.1 * d)["mean"] == 0.2


def test_flatten_season_restricted_to_a_window_is_the_window():
    from optimalinterpolation_b200.season import flatten_window, flatten_season
    obs, sie, x, y = _small_season()
    xs, ys, ts, zs = flatten_season(obs, x, y)
    for day in (0, 1, 2):
        m = (ts >= day) & (ts <= day + 8)
        xw, yw, tw, zw = flatten_window(obs, x, y, day, 9)
        assert np.array_equal(xs[m], xw) and np.array_equal(ys[m], yw) and np.array_equal(zs[m], zw)
        assert np.array_equal(ts[m] - day, tw)


@pytest.mark.gpu
def test_resident_season_equals_per_day_upload():
    """Season observations resident on the device + day window in the gather kernel == flattening and uploading every
    day's window: identical neighbour lists (after mapping the indices) and bit-identical results of both passes."""
    import optimalinterpolation_b200 as oi
    from optimalinterpolation_b200.season import run_season, flatten_season, flatten_window, day_inputs
    obs, sie, x, y = _small_season()
    a = run_season(obs, sie, x, y, days=[0, 2], radius=100, resident=True)
    b = run_season(obs, sie, x, y, days=[0, 2], radius=100, resident=False)
    assert set(a) == set(b)
    for k in b:
        if k == "_timing":
            assert set(a[k]) == set(b[k])
        elif k.endswith("_diagnostics"):
            assert np.array_equal(a[k]["nfev"], b[k]["nfev"]) and np.array_equal(a[k]["status"], b[k]["status"])
        else:
            assert np.array_equal(a[k], b[k], equal_nan=True), k
    # neighbour lists: the windowed gather returns season indices; mapped to window positions they are the window's lists
    xs, ys, ts, zs = flatten_season(obs, x, y)
    g = day_inputs(obs, sie, x, y, 2, 9)
    h = oi.Handle(0)
    h.set_observations(xs, ys, ts, zs); h.set_time_window(2, 10); h.set_cells(g["X"]); h.gather_neighbours(100e3)
    off_s, idx_s = h.get_neighbours()
    h.set_time_window(); h.set_observations(*flatten_window(obs, x, y, 2, 9)); h.set_cells(g["X"]); h.gather_neighbours(100e3)
    off_w, idx_w = h.get_neighbours()
    pos = np.cumsum((ts >= 2) & (ts <= 10)) - 1              # season index -> position inside the window
    assert np.array_equal(off_s, off_w) and np.array_equal(pos[idx_s], idx_w)
    h.close()


@pytest.mark.gpu
def test_run_season_equals_day_by_day():
    """Two days on one resident handle give exactly what two independent GPRDay runs give (handle reuse across
    days with different observation and cell counts must not leak state)."""
    import optimalinterpolation_b200 as oi
    from optimalinterpolation_b200.season import run_season, day_inputs
    from optimalinterpolation_b200.postprocess import two_pass
    obs, sie, x, y = _small_season()
    season = run_season(obs, sie, x, y, days=[0, 2], radius=100, dates=[f"201901{d:02d}" for d in range(1, 12)])
    for day in (0, 2):
        g = day_inputs(obs, sie, x, y, day, 9)
        date = f"201901{day + 5:02d}"
        gd = oi.GPRDay(g["x_train"], g["y_train"], g["t_train"], g["z"], g["X"], 100, g["mean"], g["T_mid"],
                       [np.log(25000.), np.log(25000.), 0., 0., 0., np.log(.1)])
        ref = two_pass(gd, g["ids"], g["SIE"].shape, g["SIE"], date=date)
        gd.handle.close()
        for k, v in ref.items():
            if k == "_diagnostics":
                continue
            assert np.array_equal(season[k], v, equal_nan=True), k
        assert np.isfinite(season[date + "_interp_smth"][g["ids"]]).mean() > 0.9


def _season_worker(rank, world, port, q):
    import os
    import torch.distributed as dist
    from optimalinterpolation_b200.season import run_season_sharded
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)

    def fake_runner(obs, sie, x, y, days, **kw):          # stands in for run_season (which needs a GPU)
        out = {"_timing": {}}
        for d in days:
            out[f"{d}_interp"] = np.full((3, 3), float(d)); out["_timing"][str(d)] = dict(day=d, seconds=0.1 * d, cells=9, nonfinite=0)
        return out
    a = run_season_sharded(None, None, None, None, days=range(7), collect="summary", runner=fake_runner)
    b = run_season_sharded(None, None, None, None, days=range(7), collect="all", runner=fake_runner)
    q.put((rank, sorted(k for k in a if k != "_timing"), sorted(a["_timing"]), sorted(k for k in b if k != "_timing"), float(b["5_interp"][0, 0])))
    dist.barrier(); dist.destroy_process_group()


def test_season_sharded_by_day_world2_gloo():
    """BASELINE.json configs[3]: days round-robin over ranks, one small collective for the per-day timing rows (or the
    reference's final bcast of all fields with collect='all')."""
    import socket
    import torch.multiprocessing as mp
    from optimalinterpolation_b200.season import shard_days
    assert shard_days(range(7), 0, 2) == [0, 2, 4, 6] and shard_days(range(7), 1, 2) == [1, 3, 5]
    assert sorted(sum((shard_days(range(180), r, 8) for r in range(8)), [])) == list(range(180))
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_season_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    got = {r[0]: r[1:] for r in (q.get(timeout=120) for _ in range(2))}
    [p.join(60) for p in procs]
    assert got[0][0] == ["0_interp", "2_interp", "4_interp", "6_interp"] and got[1][0] == ["1_interp", "3_interp", "5_interp"]
    assert got[0][1] == got[1][1] == [str(d) for d in range(7)]          # every rank sees every day's timing row
    assert got[0][2] == got[1][2] == sorted(f"{d}_interp" for d in range(7)) and got[0][3] == got[1][3] == 5.0
