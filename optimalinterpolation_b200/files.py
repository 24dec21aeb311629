"""On-disk contracts of the reference, so a user can feed its own file products to the GPU path
and get files back that its downstream scripts read:

* inputs  : the per-satellite season pickles written by read_and_bin.py (:16-22, :52-54: dict ``yyyymmdd -> 2-D
            gridded freeboard``, pickle protocol 2, written in < 2 GiB chunks), the sea-ice-extent pickle and the
            ``x_<res>km.npy`` / ``y_<res>km.npy`` grids (:55-57), assembled the way ``readFB`` does
            (GPR_CS2S3.py:25-63): dates present in all four streams, stacked to obs[ny, nx, 4, n_days] and
            sie[ny, nx, n_days] with concentrations below 0.15 masked out.
* outputs : the result dictionary of a day pickled with protocol 2 (GPR_CS2S3.py:192-199, :336).
No reference data files are shipped; tests/test_files.py round-trips synthetic products.
"""
from __future__ import annotations

import os
import pickle

import numpy as np

STREAMS = ("CS2_SAR", "CS2_SARIN", "S3A", "S3B")        # stream order of obs[:, :, n, :] (GPR_CS2S3.py:57)
_CHUNK = 2 ** 31 - 1


def save_pickle(dic: dict, path: str) -> None:
    """pickle protocol 2 in chunks below 2 GiB (read_and_bin.py:16-22; GPR_CS2S3.py:198-199 for results)."""
    blob = pickle.dumps(dic, protocol=2)
    with open(path, "wb") as f:
        for i in range(0, len(blob), _CHUNK):
            f.write(blob[i:i + _CHUNK])


def load_pickle(path: str) -> dict:
    with open(path, "rb") as f:
        return pickle.load(f, encoding="latin1")          # products written by Python 2 load as well


def season_paths(datapath: str, grid_res: int, season: str) -> dict:
    """File names as readFB builds them (GPR_CS2S3.py:36-45) and the grid files (:202-203)."""
    p = {s: os.path.join(datapath, f"{s}_dailyFB_{grid_res}km_{season}_season.pkl") for s in STREAMS}
    p["SIE"] = os.path.join(datapath, f"SIE_masking_{grid_res}km_{season}_season.pkl")
    p["x"] = os.path.join(datapath, f"x_{grid_res}km.npy")
    p["y"] = os.path.join(datapath, f"y_{grid_res}km.npy")
    return p


def read_season(datapath: str, grid_res: int, season: str):
    """``readFB`` (GPR_CS2S3.py:25-63) plus the grids: returns obs, sie_mask, dates, x, y."""
    paths = season_paths(datapath, grid_res, season)
    streams = [load_pickle(paths[s]) for s in STREAMS]
    sie = load_pickle(paths["SIE"])
    dates = [str(d) for d in sorted(streams[0])]
    keep = [d for d in dates if all(d in s for s in streams[1:])]
    obs = np.array([[s[d] for s in streams] for d in keep], dtype=np.float64).transpose(2, 3, 1, 0)
    sie_mask = np.array([sie[d] for d in keep], dtype=np.float64).transpose(1, 2, 0)
    with np.errstate(invalid="ignore"):
        sie_mask[sie_mask < 0.15] = np.nan
    return obs, sie_mask, keep, np.load(paths["x"]), np.load(paths["y"])


def write_season(datapath: str, grid_res: int, season: str, obs, sie, dates, x, y) -> dict:
    """Inverse of ``read_season`` (what read_and_bin.py leaves on disk): used by tests and to stage synthetic data."""
    paths = season_paths(datapath, grid_res, season)
    for n, s in enumerate(STREAMS):
        save_pickle({str(d): np.asarray(obs[:, :, n, k]) for k, d in enumerate(dates)}, paths[s])
    save_pickle({str(d): np.asarray(sie[:, :, k]) for k, d in enumerate(dates)}, paths["SIE"])
    np.save(paths["x"], x); np.save(paths["y"], y)
    return paths


def save_results(results: dict, path: str) -> None:
    """The day's result dictionary as the reference stores it (GPR_CS2S3.py:336); non-array diagnostics are dropped."""
    save_pickle({k: v for k, v in results.items() if isinstance(v, np.ndarray)}, path)
