"""Generate tests/golden/quicklook_icemask.npz from the reference's QuickLook product of 2019-01-28 (the date of the
synthetic day's seed): the packed 320x320 ice mask (cells with a freeboard), the lat/lon of the lattice corners and the
value statistics SURVEY.md Appendix D quotes.  Only where /root/reference is mounted.

    python tests/golden/make_quicklook_mask.py
"""
import os, sys
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from optimalinterpolation_b200.files import read_quicklook   # noqa: E402

SRC = "/root/reference/QuickLook Data/CS2S3_20190128_25km_quicklook.nc"
q = read_quicklook(SRC)
mask = np.isfinite(q["radar_freeboard"])
fb = q["radar_freeboard"][mask]
np.savez_compressed(os.path.join(HERE, "quicklook_icemask.npz"), source=os.path.basename(SRC), shape=np.array(mask.shape),
                    packed=np.packbits(mask), n_ice=int(mask.sum()),
                    corner_lat=np.array([q["lat"][0, 0], q["lat"][0, -1], q["lat"][-1, 0], q["lat"][-1, -1]]),
                    pole_index=np.array(np.unravel_index(np.argmax(q["lat"]), mask.shape)),
                    fb_min=float(fb.min()), fb_max=float(fb.max()), fb_mean=float(fb.mean()),
                    unc_mean=float(np.nanmean(q["uncertainty"][mask])))
print("ice cells", int(mask.sum()), "pole", np.unravel_index(np.argmax(q["lat"]), mask.shape), "fb mean", fb.mean())
