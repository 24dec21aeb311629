"""Generate tests/golden/reference_vectors.npz by running the REFERENCE'S OWN functions
(SGPkernel, SMLII, GPR3D extracted verbatim from /root/reference at run time by
oracle/reference_functions.py) on seeded synthetic cells.  Run in the build container only:

    python tests/golden/make_golden.py

The fixture pins the oracle (oracle/gpr_oracle.py) and, through it, the CUDA path, on machines
where /root/reference does not exist (the GPU box).  Library versions and the BLAS thread count (the last bits of the
LAPACK results depend on it) are stored in the file; tests/conftest.py runs the CPU suite with the same setting.
"""
import os
import sys

os.environ["OPENBLAS_NUM_THREADS"] = "1"      # before numpy loads OpenBLAS

import numpy as np
import scipy

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import warnings
warnings.simplefilter("ignore")

from optimalinterpolation_b200.synthetic import make_small_day   # noqa: E402
from oracle import reference_functions as rf                     # noqa: E402

HYPERS = {
    "x0": None,
    "notebook_optimum": np.log([2.15e5, 1.40e5, 21.0, 0.0279, 0.00346, 0.1]),
    "flat_large_ell": np.log([2.0e6, 3.0e6, 60.0, 0.5, 0.02, 0.1]),
}
CELLS = [0, 5, 100, 222, 300, 450, 600]
FIT_CELLS = [5, 100, 222, 300, 450, 512, 600]


def main():
    d = make_small_day()
    ns = rf.day_namespace(d)
    out = {"numpy": np.__version__, "scipy": scipy.__version__, "cells": np.array(CELLS), "fit_cells": np.array(FIT_CELLS)}
    for c in CELLS:
        ID = ns["X_tree"].query_ball_point(x=d.X[c, :], r=d.radius_km * 1000)
        out[f"nbr_{c}"] = np.sort(np.array(ID, dtype=np.int64))
        inputs = np.array([d.x_train[ID], d.y_train[ID], d.t_train[ID]]).T
        outputs = d.z[ID]
        mX = np.ones(len(ID)) * d.mean
        for name, h in HYPERS.items():
            h = np.array(d.x0 if h is None else h, dtype=float)
            nlZ, dnlZ = ns["SMLII"](h, inputs, outputs, mX)
            out[f"smlii_{c}_{name}"] = np.concatenate([[float(np.asarray(nlZ).reshape(-1)[0])], dnlZ])
            if c == CELLS[1]:
                ell = [np.exp(h[0]), np.exp(h[1]), np.exp(h[2])]
                K, dK = ns["SGPkernel"](inputs, grad=True, ell=ell, sigma=np.exp(h[3]))
                out[f"kernel_{name}_K"] = K[:12, :12]
                out[f"kernel_{name}_dK"] = dK[:, :12, :12]
                out[f"kernel_{name}_sums"] = np.array([K.sum(), dK[0].sum(), dK[1].sum(), dK[2].sum()])
    for c in FIT_CELLS:
        out[f"gpr3d_{c}"] = np.array(ns["GPR3D"](c), dtype=float)
    # opt=False branch with given (smoothed) hyperparameters, GPR_CS2S3.py:169-172
    ns["ellXs"] = np.tile([2.15e5, 1.40e5, 21.0], (len(d.X), 1))
    ns["sf2xs"] = np.full(len(d.X), 0.0279)
    ns["sn2xs"] = np.full(len(d.X), 0.00346)
    for c in CELLS:
        out[f"gpr3d_fixed_{c}"] = np.array(ns["GPR3D"](c, opt=False), dtype=float)
    out["blas_threads"] = os.environ["OPENBLAS_NUM_THREADS"]
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_vectors.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
