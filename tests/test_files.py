"""CPU: the reference's on-disk contracts (season pickles in, result pickle out) round-trip through files.py."""
import pickle

import numpy as np


def test_season_files_round_trip(tmp_path):
    from optimalinterpolation_b200 import files
    rng = np.random.default_rng(0)
    ny = nx = 12
    dates = ["20181101", "20181102", "20181103", "20181104"]
    obs = rng.normal(0.1, 0.05, (ny, nx, 4, 4)); obs[rng.uniform(size=obs.shape) < 0.7] = np.nan
    sie = rng.uniform(0, 1, (ny, nx, 4))
    jj, ii = np.meshgrid(np.arange(nx), np.arange(ny))
    x, y = 25000.0 * jj, 25000.0 * ii
    paths = files.write_season(str(tmp_path), 25, "2018-2019", obs, sie, dates, x, y)
    # a date missing from one stream is dropped, as in readFB (GPR_CS2S3.py:55-59)
    s3b = files.load_pickle(paths["S3B"]); del s3b["20181103"]; files.save_pickle(s3b, paths["S3B"])
    o2, m2, d2, x2, y2 = files.read_season(str(tmp_path), 25, "2018-2019")
    keep = [0, 1, 3]
    assert d2 == [dates[k] for k in keep] and o2.shape == (ny, nx, 4, 3) and m2.shape == (ny, nx, 3)
    assert np.array_equal(o2, obs[:, :, :, keep], equal_nan=True)
    ref = sie[:, :, keep].copy(); ref[ref < 0.15] = np.nan
    assert np.array_equal(m2, ref, equal_nan=True) and np.array_equal(x2, x) and np.array_equal(y2, y)
    # protocol 2, as the reference writes it (read_and_bin.py:18, GPR_CS2S3.py:199)
    raw = open(paths["S3A"], "rb").read()
    assert raw[:2] == b"\x80\x02" and set(pickle.loads(raw)) == set(dates)


def test_results_pickle(tmp_path):
    from optimalinterpolation_b200 import files
    res = {"20190128_interp": np.arange(6.0).reshape(2, 3), "20190128_ell_x_smth": np.ones((2, 3)),
           "20190128_diagnostics": {"nfev": [1, 2]}}
    p = str(tmp_path / "out.pkl")
    files.save_results(res, p)
    back = files.load_pickle(p)
    assert set(back) == {"20190128_interp", "20190128_ell_x_smth"} and np.array_equal(back["20190128_interp"], res["20190128_interp"])
    assert open(p, "rb").read()[:2] == b"\x80\x02"
