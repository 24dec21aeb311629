"""Seeded synthetic pan-Arctic day.

Produces the arrays the reference's day-setup block builds before its cell loop
(/root/reference/2021_paper_production/GPR_CS2S3.py:223-246): flattened
``x_train, y_train, t_train, z`` (stream-major, then day, then row-major lattice
order, exactly the order the reference's ``np.where``/``extend``/``concatenate``
sequence yields), the ice-cell coordinate list ``X`` and the scalar prior mean.

The lattice is what read_and_bin.py:32,49-50 produces: ``x = res*j``, ``y = res*i``
as exact doubles, so exact 300 km ties occur (GPR_CS2S3.py:159 is inclusive).
"""
from __future__ import annotations

import dataclasses
import numpy as np


@dataclasses.dataclass
class SyntheticDay:
    x_train: np.ndarray   # (n_obs,) f64 metres
    y_train: np.ndarray   # (n_obs,) f64 metres
    t_train: np.ndarray   # (n_obs,) f64 day index 0..T-1
    z: np.ndarray         # (n_obs,) f64 freeboard, metres
    X: np.ndarray         # (n_cells, 2) f64 target coords
    ids: tuple            # np.where(ice) row/col indices of the targets
    shape: tuple          # lattice shape
    mean: float           # prior mean (GPR_CS2S3.py:212)
    T: int                # days in window (GPR_CS2S3.py:206)
    T_mid: int            # prediction day index (GPR_CS2S3.py:207)
    radius_km: float      # search radius (GPR_CS2S3.py:208)
    grid_res_km: float    # lattice spacing (GPR_CS2S3.py:201)
    sat: np.ndarray | None = None  # (ny, nx, 4, T) gridded obs, NaN = none

    @property
    def x0(self) -> np.ndarray:
        """Initial log-hyperparameters, GPR_CS2S3.py:217 (six entries, last one dead)."""
        r = self.grid_res_km * 1000.0
        return np.array([np.log(r), np.log(r), np.log(1.), np.log(1.), np.log(1.), np.log(.1)])


def _truth(x, y, t, scale):
    """Smooth multi-sinusoid freeboard field, 150-400 km wavelengths (scaled with lattice)."""
    k = 2 * np.pi / (np.array([150e3, 230e3, 310e3, 400e3]) * scale)
    s = (np.sin(k[0] * x + 0.3) * np.cos(k[1] * y - 1.1)
         + np.sin(k[2] * (x + y) + 2.0) + np.cos(k[3] * (x - 0.5 * y) + 0.7)
         + 0.5 * np.sin(k[1] * x - k[2] * y))
    return 0.10 + 0.08 * (s / 3.5) + 0.003 * (t - 4.0)


def make_day(n_side: int = 320, grid_res_km: float = 25.0, ice_radius_cells: float = 78.0,
             centre: tuple = (137, 137), T: int = 9, radius_km: float = 300.0,
             tracks_per_day: int = 22, seed: int = 20190128, keep_sat: bool = False,
             s3_hole_cells: float = 38.0, cs2_hole_cells: float = 9.0,
             noise_std: float = 0.04, thin_lo: float = 0.2, ice_mask: np.ndarray | None = None) -> SyntheticDay:
    """Build one synthetic day.  Defaults are BASELINE.json configs[1] (25 km pan-Arctic day).

    ice_mask: optional (n_side, n_side) bool array that replaces the ice disc, e.g. the real ice mask of a QuickLook
    product (``files.quicklook_ice_mask``; SURVEY.md 8(d) option B): targets AND observations are confined to it."""
    res = grid_res_km * 1000.0
    jj, ii = np.meshgrid(np.arange(n_side), np.arange(n_side))
    x = res * jj.astype(np.float64)
    y = res * ii.astype(np.float64)
    ci, cj = centre
    rad = np.hypot(ii - ci, jj - cj)
    ice = rad <= ice_radius_cells
    if ice_mask is not None:
        ice = np.asarray(ice_mask, dtype=bool)
        if ice.shape != (n_side, n_side):
            raise ValueError("ice_mask must be (n_side, n_side)")
    scale = grid_res_km / 25.0
    sat = np.full((n_side, n_side, 4, T), np.nan)
    # streams: 0 CS2 SAR, 1 CS2 SARIN, 2 S3A, 3 S3B  (GPR_CS2S3.py:57)
    platforms = [(0, cs2_hole_cells), (2, s3_hole_cells), (3, s3_hole_cells)]
    for p, (stream, hole) in enumerate(platforms):
        for day in range(T):
            rng = np.random.default_rng(seed + 1000 * p + day)
            hit = np.zeros((n_side, n_side), bool)
            for _ in range(tracks_per_day):
                az = rng.uniform(0, np.pi)
                u = rng.uniform(-1, 1)   # tracks converge towards the pole
                off = ice_radius_cells * np.sign(u) * abs(u) ** 1.6
                s = np.arange(-1.5 * ice_radius_cells, 1.5 * ice_radius_cells, 0.5)
                pi_ = np.rint(ci + off * np.cos(az) + s * np.sin(az)).astype(int)
                pj_ = np.rint(cj - off * np.sin(az) + s * np.cos(az)).astype(int)
                ok = (pi_ >= 0) & (pi_ < n_side) & (pj_ >= 0) & (pj_ < n_side)
                hit[pi_[ok], pj_[ok]] = True
            hit &= ice & (rad > hole)
            # smooth regional data drop-out (quality filtering), widens the n histogram
            keep = thin_lo + (1 - thin_lo) * (0.5 + 0.5 * np.sin(2 * np.pi * x / (2400e3 * scale) + 0.5)
                                              * np.cos(2 * np.pi * y / (3100e3 * scale) - 0.4))
            hit &= rng.uniform(size=hit.shape) < keep
            idx = np.where(hit)
            f = _truth(x[idx], y[idx], float(day), scale)
            val = np.clip(f + rng.normal(0.0, noise_std, size=f.shape), -0.37, 0.63)
            if stream == 0:
                # CS2 switches to SARIn mode in an outer (coastal) ring
                sarin = rad[idx] > 0.8 * ice_radius_cells
                sat[idx[0][~sarin], idx[1][~sarin], 0, day] = val[~sarin]
                sat[idx[0][sarin], idx[1][sarin], 1, day] = val[sarin]
            else:
                sat[idx[0], idx[1], stream, day] = val
    xs, ys, ts, zs = [], [], [], []
    for stream in range(4):          # stream-major (GPR_CS2S3.py:238-241)
        for day in range(T):         # then day (GPR_CS2S3.py:227)
            idx = np.where(~np.isnan(sat[:, :, stream, day]))  # row-major
            xs.append(x[idx]); ys.append(y[idx])
            ts.append(np.ones(idx[0].size) * day)
            zs.append(sat[:, :, stream, day][idx])
    x_train = np.concatenate(xs); y_train = np.concatenate(ys)
    t_train = np.concatenate(ts); z = np.concatenate(zs)
    ids = np.where(ice)
    X = np.array([x[ids], y[ids]]).T.copy()
    mean = float(np.round(np.mean(z), 3))
    return SyntheticDay(x_train, y_train, t_train, z, X, ids, (n_side, n_side), mean, T, T // 2,
                        radius_km, grid_res_km, sat if keep_sat else None)


def make_small_day(seed: int = 7, n_side: int = 40, ice_radius_cells: float = 14.0,
                   radius_km: float = 100.0, tracks_per_day: int = 4) -> SyntheticDay:
    """A scaled-down day (n per cell ~ 40-300) that the CPU oracle fits in seconds."""
    return make_day(n_side=n_side, ice_radius_cells=ice_radius_cells,
                    centre=(n_side // 2, n_side // 2), radius_km=radius_km,
                    tracks_per_day=tracks_per_day, seed=seed,
                    s3_hole_cells=4.0, cs2_hole_cells=1.0)


def make_day_cfg5(seed: int = 20190128, tracks_per_day: int = 17) -> SyntheticDay:
    """BASELINE.json configs[4]: the 12.5 km lattice with a 500 km search radius (the reference's knobs
    ``grid_res = 12.5`` at GPR_CS2S3.py:201 with 640x640 product grids, ``radius = 500`` at :208); same ice disc, holes
    and track geometry as ``make_day`` in kilometres.  ~76k ice cells, ~60k observations, n per cell ~1500...5100."""
    return make_day(n_side=640, grid_res_km=12.5, ice_radius_cells=156.0, centre=(274, 274), radius_km=500.0,
                    tracks_per_day=tracks_per_day, seed=seed, s3_hole_cells=76.0, cs2_hole_cells=18.0)


def make_day_real_mask(mask_npz: str, **kw) -> SyntheticDay:
    """SURVEY.md 8(d) option B: the synthetic tracks of ``make_day`` over the REAL ice mask of a QuickLook day
    (tests/golden/quicklook_icemask.npz, written by tests/golden/make_quicklook_mask.py; 17 697 ice cells on
    2019-01-28, pole at lattice index (137, 137))."""
    f = np.load(mask_npz)
    shape = tuple(int(v) for v in f["shape"])
    mask = np.unpackbits(f["packed"])[:shape[0] * shape[1]].reshape(shape).astype(bool)
    ci, cj = (int(v) for v in f["pole_index"])
    # track geometry: 34 tracks per platform and day within 110 cells of the pole (the mask reaches 137 cells out): n per
    # cell 2 ... 2326, median 723, sum n^3 = 1.04x the disc day's -- ragged down to a handful of observations at the ice edge
    kw.setdefault("ice_radius_cells", 110.0); kw.setdefault("tracks_per_day", 34)
    return make_day(n_side=shape[0], centre=(ci, cj), ice_mask=mask, **kw)
