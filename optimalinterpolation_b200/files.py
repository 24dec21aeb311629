"""On-disk contracts of the reference, so a user can feed its own file products to the GPU path
and get files back that its downstream scripts read:

* inputs  : the per-satellite season pickles written by read_and_bin.py (:16-22, :52-54: dict ``yyyymmdd -> 2-D
            gridded freeboard``, pickle protocol 2, written in < 2 GiB chunks), the sea-ice-extent pickle and the
            ``x_<res>km.npy`` / ``y_<res>km.npy`` grids (:55-57), assembled the way ``readFB`` does
            (GPR_CS2S3.py:25-63): dates present in all four streams, stacked to obs[ny, nx, 4, n_days] and
            sie[ny, nx, n_days] with concentrations below 0.15 masked out.
* outputs : the result dictionary of a day pickled with protocol 2 (GPR_CS2S3.py:192-199, :336).
* QuickLook: the 232 ``CS2S3_yyyymmdd_25km_quicklook.nc`` products under the reference's ``QuickLook Data/`` (daily
            320x320 ``lat, lon, radar_freeboard, uncertainty``).  They are netCDF-4/HDF5 files whose four
            variables are stored uncompressed and contiguous, so ``read_quicklook`` reads them by offset without an HDF5
            library (none is installed) after checking the signature and the size.  Their real ice masks are the second
            bench geometry (``quicklook_ice_mask``; tests/golden/quicklook_icemask.npz holds one packed mask).
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


# ---- QuickLook netCDF-4 products -----------------------------------------------------------
QUICKLOOK_SHAPE = (320, 320)
QUICKLOOK_OFFSET = 8232                                   # first byte of the contiguous data block
QUICKLOOK_BYTES = 3285032                                 # every file of the reference's QuickLook Data/
QUICKLOOK_VARS = ("lat", "lon", "radar_freeboard", "uncertainty")
_HDF5_SIGNATURE = b"\x89HDF\r\n\x1a\n"


def read_quicklook(path: str) -> dict:
    """One QuickLook day: dict of four (320, 320) float64 arrays (NaN off the ice).  Raises ValueError when the file is
    not laid out like the reference's products (wrong signature or size), rather than returning garbage."""
    size = os.path.getsize(path)
    with open(path, "rb") as f:
        sig = f.read(8)
    if sig != _HDF5_SIGNATURE:
        raise ValueError(f"{path}: not an HDF5/netCDF-4 file")
    n = QUICKLOOK_SHAPE[0] * QUICKLOOK_SHAPE[1]
    if size != QUICKLOOK_BYTES or size < QUICKLOOK_OFFSET + 4 * n * 8:
        raise ValueError(f"{path}: {size} bytes, expected the QuickLook layout of {QUICKLOOK_BYTES} bytes "
                         "(four contiguous <f8 320x320 arrays at offset 8232)")
    a = np.fromfile(path, dtype="<f8", offset=QUICKLOOK_OFFSET, count=4 * n).reshape((4,) + QUICKLOOK_SHAPE)
    out = {k: a[i].astype(np.float64) for i, k in enumerate(QUICKLOOK_VARS)}
    lat = out["lat"]
    if not (np.isfinite(lat).all() and 30.0 < lat.min() and lat.max() <= 90.0):
        raise ValueError(f"{path}: latitude block out of range -- not the QuickLook layout")
    return out


def quicklook_ice_mask(path: str) -> np.ndarray:
    """(320, 320) bool: cells that carry a freeboard in that day's product = the day's ice cells (the set GPR_CS2S3.py:243
    takes from the SIE mask)."""
    return np.isfinite(read_quicklook(path)["radar_freeboard"])


def write_quicklook_like(path: str, lat, lon, fb, unc) -> None:
    """A file with the QuickLook byte layout (HDF5 signature, zero-filled header block, the four arrays): lets the reader
    be tested where the reference's data are not mounted.  It is NOT a valid netCDF file."""
    blob = bytearray(QUICKLOOK_BYTES)
    blob[:8] = _HDF5_SIGNATURE
    data = np.stack([np.asarray(v, dtype="<f8").reshape(QUICKLOOK_SHAPE) for v in (lat, lon, fb, unc)]).tobytes()
    blob[QUICKLOOK_OFFSET:QUICKLOOK_OFFSET + len(data)] = data
    with open(path, "wb") as f:
        f.write(bytes(blob))
