"""Run the reference's own functions, unmodified, straight from /root/reference.

TEST INFRASTRUCTURE ONLY.  /root/reference/2021_paper_production/GPR_CS2S3.py cannot be
imported as a module (syntax error at :317, Python-2 ``izip_longest`` at :270/:325, mpi4py
and astropy imports, module-level loads of files that are not shipped at :202-210), so the
three hot-path functions are pulled out by ``ast`` at run time and exec'd with the module
globals they read.  No reference source text is stored in this repository.

Only usable where /root/reference exists (the build container); the GPU box uses the
numpy restatement in ``oracle/gpr_oracle.py`` plus the committed golden vectors.
"""
from __future__ import annotations

import ast
import os

import numpy as np
import scipy
import scipy.optimize
import scipy.spatial
from scipy.spatial.distance import cdist, pdist, squareform

REFERENCE_FILE = "/root/reference/2021_paper_production/GPR_CS2S3.py"
HOT_PATH_FUNCTIONS = ("SGPkernel", "SMLII", "GPR3D")   # GPR_CS2S3.py:78, :107, :143


def available() -> bool:
    return os.path.exists(REFERENCE_FILE)


def _function_sources(path: str = REFERENCE_FILE) -> dict:
    text = open(path).read()
    lines = text.splitlines(keepends=True)
    # ast.parse fails on the whole file (:317), so locate each def by its header line and parse
    # that block on its own.
    out = {}
    for name in HOT_PATH_FUNCTIONS:
        start = next(i for i, l in enumerate(lines) if l.startswith(f"def {name}("))
        end = start + 1
        while end < len(lines) and (lines[end].startswith((" ", "\t")) or not lines[end].strip()):
            end += 1
        block = "".join(lines[start:end])
        tree = ast.parse(block)
        out[name] = ast.get_source_segment(block, tree.body[0])
    return out


def load(day_globals: dict | None = None) -> dict:
    """Return a namespace holding the reference's SGPkernel/SMLII/GPR3D.

    ``day_globals`` supplies the module-level names GPR3D reads (GPR_CS2S3.py:201-217,
    :238-246, :313-315): X_tree, X, x_train, y_train, t_train, z, radius, mean, T_mid, x0,
    and for opt=False ellXs, sf2xs, sn2xs.
    """
    ns = {"np": np, "scipy": scipy, "squareform": squareform, "pdist": pdist, "cdist": cdist}
    if day_globals:
        ns.update(day_globals)
    for name, src in _function_sources().items():
        exec(compile(src, f"{REFERENCE_FILE}::{name}", "exec"), ns)
    return ns


def day_namespace(day, x0=None) -> dict:
    """Reference namespace wired to a SyntheticDay (mirrors GPR_CS2S3.py:238-246)."""
    xy_train = np.array([day.x_train, day.y_train]).T
    g = dict(X_tree=scipy.spatial.cKDTree(xy_train), X=day.X, x_train=day.x_train,
             y_train=day.y_train, t_train=day.t_train, z=day.z, radius=day.radius_km,
             mean=day.mean, T_mid=day.T_mid, x0=list(day.x0 if x0 is None else x0))
    return load(g)
