"""CPU: the host-side smoothing/assembly that mirrors GPR_CS2S3.py:65-76, :282-307."""
import numpy as np

from optimalinterpolation_b200.postprocess import assemble, gaussian2d_kernel, nan_convolve, smooth


def _brute(data, kernel):
    """astropy's documented algorithm, pixel by pixel: interpolate over NaN, zero fill outside, normalised kernel."""
    k = kernel / kernel.sum()
    h = k.shape[0] // 2
    out = np.full(data.shape, np.nan)
    for i in range(data.shape[0]):
        for j in range(data.shape[1]):
            top = bot = 0.0
            for a in range(-h, h + 1):
                for b in range(-h, h + 1):
                    ii, jj = i + a, j + b
                    w = k[h - a, h - b]
                    if 0 <= ii < data.shape[0] and 0 <= jj < data.shape[1]:
                        v = data[ii, jj]
                        if not np.isnan(v):
                            top += w * v; bot += w
                    else:
                        bot += w          # boundary='fill', fill_value=0
            if bot != 0:
                out[i, j] = top / bot
    return out


def test_kernel_shape_and_symmetry():
    for std, size in ((1, 9), (2, 17)):
        k = gaussian2d_kernel(std)
        assert k.shape == (size, size) and np.allclose(k, k.T) and np.allclose(k, k[::-1, ::-1])
        assert np.isclose(k[size // 2, size // 2], 1 / (2 * np.pi * std ** 2))


def test_nan_convolve_matches_pixel_loop():
    rng = np.random.default_rng(0)
    d = rng.normal(size=(23, 19))
    d[rng.uniform(size=d.shape) < 0.3] = np.nan
    d[:6, :7] = np.nan
    for std in (1, 2):
        k = gaussian2d_kernel(std)
        assert np.allclose(nan_convolve(d, k), _brute(d, k), rtol=1e-12, atol=1e-14, equal_nan=True)


def test_smooth_semantics():
    shape = (40, 40)
    mask = np.full(shape, np.nan); mask[10:30, 10:30] = 1.0
    data = np.full(shape, np.nan); data[10:30, 10:30] = 5.0
    data[15, 15] = np.inf; data[20, 20] = 1e9; data[12, 12] = np.nan
    s = smooth(data, vmax=7.0, mask=mask, std=2)
    assert np.isnan(s[~(mask == 1.0)]).all() and np.isfinite(s[10:30, 10:30]).all()
    assert abs(s[12, 12] - 5.0) < 1e-3 and abs(s[15, 15] - 5.0) < 1e-3    # NaN/inf holes are interpolated over
    assert 5.0 < s[20, 20] < 7.0                                            # clipped to vmax before smoothing
    const = np.full(shape, np.nan); const[10:30, 10:30] = 3.0
    assert np.allclose(smooth(const, 10.0, mask, 1)[10:30, 10:30], 3.0)


def test_assemble_keys():
    ids = (np.array([0, 1, 2]), np.array([2, 1, 0]))
    out = np.arange(24, dtype=float).reshape(3, 8)
    res = assemble(out, ids, (3, 3), "20190128")
    assert set(res) == {"20190128" + s for s in ("_interp", "_interp_error", "_lZ", "_ell_x", "_ell_y", "_ell_t", "_sf2", "_sn2")}
    assert res["20190128_lZ"][1, 1] == 10.0 and np.isnan(res["20190128_lZ"][0, 0])
