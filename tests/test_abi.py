"""CPU: the C-ABI library loads and exports every symbol include/oi_b200.h declares; argument
errors are reported without a GPU; there is no CPU fallback."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from optimalinterpolation_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    return _lib.load()


def test_header_symbols_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "oi_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(oi_[a-z_]+)\s*\(", hdr))
    assert len(names) >= 13
    for n in names:
        assert hasattr(lib, n), n
    from optimalinterpolation_b200 import _lib
    assert names == set(_lib.EXPORTS)


def test_struct_layout_matches_header():
    import ctypes as C
    from optimalinterpolation_b200 import _lib
    assert C.sizeof(_lib.OiParams) == 3 * 8 + 4 * 4 + 6 * 8 + 2 * 8 + 2 * 4
    assert C.sizeof(_lib.OiStats) == 9 * 8 + 7 * 8 + 3 * 8 + 3 * 8


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import optimalinterpolation_b200 as oi
    with pytest.raises(oi.OIError, match="no CPU fallback"):
        oi.Handle(0)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "optimalinterpolation_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f
