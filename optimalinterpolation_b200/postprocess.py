"""Host-side post-processing of a day's results, mirroring GPR_CS2S3.py:264-336.

* ``assemble``  : scatter per-cell result rows into 2-D NaN grids (``fs_grid[IDs] = ...``, :282-297)
* ``smooth``    : the reference's hyperparameter smoothing (:65-76) -- inf -> NaN, clip to vmax, NaN-aware
                  normalised Gaussian convolution, exact zeros -> nanmean, re-mask.  The reference calls
                  ``astropy.convolution.convolve(data, Gaussian2DKernel(x_stddev=std, y_stddev=std))`` with its
                  defaults (boundary='fill', fill_value=0, nan_treatment='interpolate', normalize_kernel=True,
                  kernel size 8*std+1).  astropy is a third-party dependency that is not installed here; its
                  published algorithm is restated: out = sum(k * v over non-NaN pixels) / sum(k over non-NaN
                  pixels), pixels outside the array counting as (non-NaN) zeros.
* ``two_pass``  : pass 1 (fit + predict) -> smooth the five hyperparameter fields (:299-307) -> pass 2
                  (predict with the smoothed fields, :311-320), returning the reference's result dict (:290-297,
                  :303-307, :333-334).
These are O(grid) numpy operations on a 320x320 field; they stay on the host.
"""
from __future__ import annotations

import numpy as np
from scipy.signal import convolve2d


def gaussian2d_kernel(std: float) -> np.ndarray:
    """astropy ``Gaussian2DKernel(x_stddev=std, y_stddev=std)``: size round-up-to-odd(8*std), 'center'
    discretisation of amplitude 1/(2 pi std^2) * exp(-(x^2+y^2)/(2 std^2))."""
    size = int(np.ceil(8 * std))
    if size % 2 == 0:
        size += 1
    r = np.arange(size) - size // 2
    xx, yy = np.meshgrid(r, r)
    return np.exp(-(xx ** 2 + yy ** 2) / (2.0 * std ** 2)) / (2 * np.pi * std ** 2)


def nan_convolve(data: np.ndarray, kernel: np.ndarray) -> np.ndarray:
    """convolve(..., boundary='fill', fill_value=0, nan_treatment='interpolate', normalize_kernel=True)."""
    k = kernel / kernel.sum()
    nan = np.isnan(data)
    top = convolve2d(np.where(nan, 0.0, data), k, mode="same", boundary="fill", fillvalue=0.0)
    bot = convolve2d((~nan).astype(float), k, mode="same", boundary="fill", fillvalue=1.0)
    with np.errstate(invalid="ignore", divide="ignore"):
        out = top / bot
    out[bot == 0] = np.nan
    return out


def smooth(data, vmax, mask, std=1):
    """GPR_CS2S3.py:65-76."""
    data_smth = np.copy(data)
    data_smth[np.isinf(data_smth)] = np.nan
    with np.errstate(invalid="ignore"):
        data_smth[data_smth > vmax] = vmax
    data_smth = nan_convolve(data_smth, gaussian2d_kernel(std))
    data_smth[data_smth == 0] = np.nanmean(data_smth)
    data_smth[np.isnan(mask)] = np.nan
    return data_smth


def assemble(out: np.ndarray, ids, shape, date: str = "") -> dict:
    """Result rows (n_cells, 8) -> the eight 2-D fields of GPR_CS2S3.py:282-297."""
    names = ("_interp", "_interp_error", "_lZ", "_ell_x", "_ell_y", "_ell_t", "_sf2", "_sn2")
    res = {}
    for k, nm in enumerate(names):
        g = np.zeros(shape) * np.nan
        g[ids] = out[:, k]
        res[date + nm] = g
    return res


def two_pass(gpr, ids, shape, sie_mask, date: str = "", grid_res: float = 25, T: int = 9, **run_kw) -> dict:
    """The reference's whole day product on a ``GPRDay``: pass 1, smoothing, pass 2."""
    r1 = gpr.run(opt=True, **run_kw)
    res = assemble(r1["out"], ids, shape, date)
    std = 2 if grid_res == 25 else 1                                      # :299-302
    radius = gpr.radius
    res[date + "_ell_x_smth"] = smooth(res[date + "_ell_x"], 2 * radius * 1000, sie_mask, std)   # :303-307
    res[date + "_ell_y_smth"] = smooth(res[date + "_ell_y"], 2 * radius * 1000, sie_mask, std)
    res[date + "_ell_t_smth"] = smooth(res[date + "_ell_t"], T, sie_mask, std)
    res[date + "_sf2_smth"] = smooth(res[date + "_sf2"], 0.1, sie_mask, std)
    res[date + "_sn2_smth"] = smooth(res[date + "_sn2"], 0.05, sie_mask, std)
    ellXs = np.array([res[date + "_ell_x_smth"][ids], res[date + "_ell_y_smth"][ids], res[date + "_ell_t_smth"][ids]]).T
    sn2xs = res[date + "_sn2_smth"][ids]                                   # :313-315
    sf2xs = res[date + "_sf2_smth"][ids]
    r2 = gpr.run(opt=False, ellXs=ellXs, sf2xs=sf2xs, sn2xs=sn2xs)
    fs_smth = np.zeros(shape) * np.nan
    sfs2_smth = np.zeros(shape) * np.nan
    fs_smth[ids] = r2["out"][:, 0]
    sfs2_smth[ids] = r2["out"][:, 1]
    res[date + "_interp_smth"] = fs_smth                                   # :333-334
    res[date + "_interp_error_smth"] = sfs2_smth
    res["_diagnostics"] = dict(n=r1["n"], nfev=r1["nfev"], status=r1["status"], status_smth=r2["status"])
    return res
