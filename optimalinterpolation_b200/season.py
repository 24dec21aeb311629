"""Day setup and season loop around the batched GP path, mirroring GPR_CS2S3.py:201-246.

The reference script processes ONE hard-coded day (``day = 1``, :211): it slices the ``T``-day window out of the
season's gridded observations (:213), flattens it stream-major / day-major / row-major into ``x_train, y_train,
t_train, z`` (:223-241), takes the ice cells of the middle day (:214, :243-244) and a prior mean (:212), and then
runs both passes.  ``run_season`` repeats exactly that setup for a list of days on one resident GPU handle.
Everything here is O(grid) host work.  Two ways to feed the GPU: per day (the window is flattened on the host and
uploaded by ``oi_gpr_day``, ~1 MB) or ``resident=True``: the whole season is flattened once (``flatten_season``, same
stream-major / day-major / row-major order with the absolute day index as t), stays on the device, and each day only
moves the gather's day window (``oi_set_time_window``) - the neighbour lists, their order and the window-relative
time coordinate are the same, so both ways give bit-identical results.
"""
from __future__ import annotations

import numpy as np

from .gpr import GPRDay, Handle
from .postprocess import assemble, two_pass


def flatten_window(obs: np.ndarray, x: np.ndarray, y: np.ndarray, day: int, T: int = 9):
    """``sat = obs[:, :, :, day:day+T]`` (:213) flattened as at :223-241.

    obs: (ny, nx, n_streams, n_days) gridded observations, NaN where empty; x, y: (ny, nx) cell coordinates.
    Returns x_train, y_train, t_train (day index inside the window, 0..T-1), z.
    """
    sat = obs[:, :, :, day:day + T]
    xs, ys, ts, zs = [], [], [], []
    for stream in range(sat.shape[2]):            # x1.., x2.., x3.., x4.. concatenated (:238-241)
        for d in range(sat.shape[3]):             # for day in range(sat.shape[3]) (:227)
            field = sat[:, :, stream, d]
            idx = np.where(~np.isnan(field))      # row-major (:228-231)
            xs.append(x[idx]); ys.append(y[idx])
            ts.append(np.ones(idx[0].size) * d)
            zs.append(field[idx])
    return (np.concatenate(xs).astype(np.float64), np.concatenate(ys).astype(np.float64),
            np.concatenate(ts).astype(np.float64), np.concatenate(zs).astype(np.float64))


def flatten_season(obs: np.ndarray, x: np.ndarray, y: np.ndarray):
    """All days of ``obs`` flattened like a window (:223-241) with t = absolute day index: restricting it to
    day <= t <= day+T-1 leaves exactly ``flatten_window(obs, x, y, day, T)`` (same order, t shifted by ``day``)."""
    return flatten_window(obs, x, y, 0, obs.shape[3])


class _ResidentDay:
    """``GPRDay``-like front end (what ``postprocess.two_pass`` needs) on a handle that already holds the season's
    observations: selects the day window, uploads the day's cells and gathers once, then runs either pass."""

    def __init__(self, handle: Handle, day: int, T: int, X, radius, mean, T_mid, x0):
        self.handle, self.radius, self.mean, self.T_mid, self.x0 = handle, float(radius), float(mean), float(T_mid), list(x0)
        handle.set_time_window(day, day + T - 1)
        handle.set_cells(X)
        handle.gather_neighbours(self.radius * 1000.0)

    def run(self, opt: bool = True, ellXs=None, sf2xs=None, sn2xs=None, **kw):
        hin = None if opt else np.column_stack([np.asarray(ellXs, float), np.asarray(sf2xs, float), np.asarray(sn2xs, float)])
        p = self.handle.make_params(self.radius * 1000.0, self.T_mid, self.mean, self.x0, mode=0 if opt else 1, **kw)
        self.handle.run(p, hin)
        return self.handle.get_results()


def day_inputs(obs, sie_mask, x, y, day: int, T: int = 9, prior_mean=None) -> dict:
    """The module globals of GPR_CS2S3.py:207-246 for one day index (the interpolated day is day + T//2).

    prior_mean: float, or callable(day) -> float.  The reference uses the mean of a separate CS2 first-year-ice
    product over the previous 9 days (:212, data not shipped); the default here is the rounded mean of the
    window's own observations, as in the synthetic generator.
    """
    T_mid = T // 2
    x_train, y_train, t_train, z = flatten_window(obs, x, y, day, T)
    SIE = sie_mask[:, :, day + T_mid]                                    # :214
    ids = np.where(~np.isnan(SIE))                                       # :243
    X = np.array([x[ids], y[ids]]).T.astype(np.float64).copy()           # :244
    if prior_mean is None:
        mean = float(np.round(np.mean(z), 3)) if z.size else 0.0
    elif callable(prior_mean):
        mean = float(prior_mean(day))
    else:
        mean = float(prior_mean)
    return dict(x_train=x_train, y_train=y_train, t_train=t_train, z=z, X=X, ids=ids, SIE=SIE, mean=mean,
                T=T, T_mid=T_mid)


def run_season(obs, sie_mask, x, y, days, dates=None, grid_res: float = 25, T: int = 9, radius: float = 300,
               x0=None, prior_mean=None, smooth_pass: bool = True, device: int = 0, handle: Handle | None = None,
               resident: bool = False, **run_kw) -> dict:
    """Both passes (or pass 1 only) for every day index in ``days`` on one GPU handle.

    Returns one dict with the reference's per-date keys (``<date>_interp``, ``<date>_ell_x_smth``, ...,
    GPR_CS2S3.py:290-297, :303-307, :333-334); ``dates[day + T//2]`` names a day (default: the index)."""
    if x0 is None:
        x0 = [np.log(grid_res * 1000), np.log(grid_res * 1000), np.log(1.), np.log(1.), np.log(1.), np.log(.1)]   # :217
    import time
    own = handle is None
    handle = handle or Handle(device)
    out = {}
    timing = out.setdefault("_timing", {})        # per date: seconds of both passes, cells, non-finite cells
    try:
        if resident:
            handle.set_observations(*flatten_season(obs, x, y))          # once; every day only moves the window
        for day in days:
            g = day_inputs(obs, sie_mask, x, y, day, T, prior_mean)
            date = str(dates[day + g["T_mid"]]) if dates is not None else str(day + g["T_mid"])
            if resident:
                gd = _ResidentDay(handle, day, T, g["X"], radius, g["mean"], g["T_mid"], x0)
            else:
                gd = GPRDay(g["x_train"], g["y_train"], g["t_train"], g["z"], g["X"], radius, g["mean"], g["T_mid"], x0,
                            handle=handle)
            t0 = time.perf_counter()
            if smooth_pass:
                res = two_pass(gd, g["ids"], g["SIE"].shape, g["SIE"], date=date, grid_res=grid_res, T=T, **run_kw)
                res[date + "_diagnostics"] = res.pop("_diagnostics")
            else:
                r1 = gd.run(opt=True, **run_kw)
                res = assemble(r1["out"], g["ids"], g["SIE"].shape, date)
            timing[date] = dict(day=int(day), seconds=time.perf_counter() - t0, cells=int(len(g["X"])),
                                nonfinite=int(np.isnan(res[date + "_interp"][g["ids"]]).sum()))
            out.update(res)
    finally:
        if resident and not own:
            handle.set_time_window()                                      # leave a borrowed handle without a window
        if own:
            handle.close()
    return out


# ------------------------------------------------------------------------------------------------------------------
# BASELINE.json configs[3]: the winter season (~180 daily fields) on the 8 GPUs of one box.  Days are independent
# (the reference script IS one day, GPR_CS2S3.py:211; the season is that script run per day), so the natural shard is
# BY DAY: every rank holds the whole season's gridded observations (as every MPI rank of the reference loads all data,
# :201-210), takes days[rank::world] and runs both passes for them on its own GPU.  No cost model is needed (a day is a
# day), there is no per-day cross-rank tail, and the only exchange is the final collection of the per-day products.
# ------------------------------------------------------------------------------------------------------------------
def shard_days(days, rank: int, world: int) -> list:
    """Round-robin: consecutive days (similar ice extent, similar cost) land on different ranks."""
    return list(days)[rank::world]


def run_season_sharded(obs, sie_mask, x, y, days, rank: int | None = None, world: int | None = None,
                       collect: str = "summary", runner=None, **kw) -> dict:
    """``run_season`` for this rank's share of ``days`` (one process per GPU, launched with torchrun).

    collect: "summary" -- every rank returns its own days' fields plus, under ``_timing``, the per-day timing rows of
             ALL ranks (one small all_gather_object); "all" -- additionally every rank receives every day's fields
             (the reference's final ``COMM.bcast`` of the day dictionary, :311); "none" -- no collective at all.
    Without an initialised process group it is ``run_season`` on one GPU.  ``runner`` (default ``run_season``) is the
    per-rank worker; the CPU tests substitute it."""
    import torch.distributed as dist
    ddp = dist.is_available() and dist.is_initialized()
    if rank is None:
        rank = dist.get_rank() if ddp else 0
    if world is None:
        world = dist.get_world_size() if ddp else 1
    mine = shard_days(days, rank, world)
    runner = runner or run_season
    out = runner(obs, sie_mask, x, y, mine, **kw) if mine else {"_timing": {}}
    if ddp and world > 1 and collect != "none":
        if collect == "all":
            parts = [None] * world
            dist.all_gather_object(parts, out)
            merged = {"_timing": {}}
            for p in parts:
                merged["_timing"].update(p.pop("_timing"))
                merged.update(p)
            return merged
        rows = [None] * world
        dist.all_gather_object(rows, out["_timing"])
        out["_timing"] = {k: v for r in rows for k, v in r.items()}
    return out


def make_synthetic_season(n_days: int, T: int = 9, **day_kw):
    """Gridded observations, ice mask and lattice coordinates of a synthetic season with ``n_days`` interpolated days
    (n_days + T - 1 days of tracks), in the layout ``readFB`` returns (GPR_CS2S3.py:25-63): obs (ny, nx, 4, days),
    SIE (ny, nx, days) NaN off-ice, x / y (ny, nx)."""
    from .synthetic import make_day
    d = make_day(T=n_days + T - 1, keep_sat=True, **day_kw)
    ny, nx = d.shape
    res = d.grid_res_km * 1000.0
    jj, ii = np.meshgrid(np.arange(nx), np.arange(ny))
    sie = np.full((ny, nx, n_days + T - 1), np.nan)
    sie[d.ids[0], d.ids[1], :] = 1.0
    return d.sat, sie, res * jj.astype(np.float64), res * ii.astype(np.float64)
