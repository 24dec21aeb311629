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


def test_quicklook_reader(tmp_path):
    """QuickLook products are read by offset; anything that is not laid out like them is refused."""
    import os
    import pytest
    from optimalinterpolation_b200 import files
    rng = np.random.default_rng(1)
    lat = rng.uniform(40, 89.9, files.QUICKLOOK_SHAPE); lon = rng.uniform(-180, 180, files.QUICKLOOK_SHAPE)
    fb = rng.normal(0.1, 0.05, files.QUICKLOOK_SHAPE); fb[rng.uniform(size=fb.shape) < 0.8] = np.nan
    unc = np.where(np.isnan(fb), np.nan, 0.02)
    p = str(tmp_path / "CS2S3_20190128_25km_quicklook.nc")
    files.write_quicklook_like(p, lat, lon, fb, unc)
    q = files.read_quicklook(p)
    assert np.array_equal(q["lat"], lat) and np.array_equal(q["radar_freeboard"], fb, equal_nan=True)
    assert np.array_equal(files.quicklook_ice_mask(p), np.isfinite(fb))
    bad = str(tmp_path / "short.nc")
    open(bad, "wb").write(open(p, "rb").read()[:100000])
    with pytest.raises(ValueError):
        files.read_quicklook(bad)
    notnc = str(tmp_path / "x.nc")
    open(notnc, "wb").write(b"\0" * files.QUICKLOOK_BYTES)
    with pytest.raises(ValueError):
        files.read_quicklook(notnc)
    # the reference's own files, where they are mounted, against SURVEY.md Appendix D and the committed mask fixture
    src = "/root/reference/QuickLook Data/CS2S3_20190128_25km_quicklook.nc"
    fix = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "quicklook_icemask.npz"))
    assert int(fix["n_ice"]) == 17697 and tuple(fix["pole_index"]) == (137, 137)
    if os.path.exists(src):
        q = files.read_quicklook(src)
        assert 36.0 < q["lat"].min() < 37.0 and 89.8 < q["lat"].max() < 90.0 and -180 <= q["lon"].min() and q["lon"].max() <= 180
        m = np.isfinite(q["radar_freeboard"])
        assert m.sum() == 17697 and np.array_equal(np.packbits(m), fix["packed"])
        assert -0.3 < np.nanmin(q["radar_freeboard"]) and np.nanmax(q["radar_freeboard"]) < 0.6


def test_real_mask_day_geometry():
    import os
    from optimalinterpolation_b200.synthetic import make_day_real_mask
    d = make_day_real_mask(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "quicklook_icemask.npz"))
    assert len(d.X) == 17697 and d.shape == (320, 320)
    from scipy.spatial import cKDTree
    cnt = np.asarray(cKDTree(np.c_[d.x_train, d.y_train]).query_ball_point(d.X[::50], r=300e3, return_length=True))
    assert cnt.min() >= 1 and 400 < np.median(cnt) < 1500 and len(d.z) == 40424
