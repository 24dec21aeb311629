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
    L = C.CDLL(_lib.LIB_PATH)
    # the library reports sizeof() of its own structs; field order is checked against the header text
    assert C.sizeof(_lib.OiParams) == L.oi_sizeof_params()
    assert C.sizeof(_lib.OiStats) == L.oi_sizeof_stats()
    hdr = open(os.path.join(ROOT, "include", "oi_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    for cname, cls in (("oi_params", _lib.OiParams), ("oi_stats", _lib.OiStats)):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (cname, cname), hdr, flags=re.S).group(1)
        fields = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            names = decl.split(None, 1)[1]
            fields += [re.sub(r"\[.*?\]", "", n).strip() for n in names.split(",")]
        assert fields == [f for f, _ in cls._fields_], cname


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
            if f.endswith((".py", ".cu", ".h", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f


def test_header_is_plain_c_and_links(tmp_path):
    """include/oi_b200.h must compile as C (no C++/torch types in the boundary) and a C program linked against the
    shared library must run the calls that need no GPU."""
    import subprocess
    from optimalinterpolation_b200 import _lib
    src = tmp_path / "abi.c"
    src.write_text('#include <stdio.h>\n#include "oi_b200.h"\n'
                   'int main(void) {\n'
                   '  oi_params p = {0}; oi_stats s; (void)s; p.mode = OI_MODE_FIT; p.engine = OI_ENGINE_LOCKSTEP;\n'
                   '  printf("%d %d %d\\n", oi_version(), oi_sizeof_params() == (int)sizeof(oi_params), oi_sizeof_stats() == (int)sizeof(oi_stats));\n'
                   '  return oi_run(NULL, &p, NULL) == OI_ERR_ARG ? 0 : 1;\n}\n')
    exe = tmp_path / "abi"
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                           "-L", libdir, "-loi_b200", "-Wl,-rpath," + libdir])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    ver, okp, oks = out.stdout.split()
    assert int(ver) >= 110 and okp == "1" and oks == "1"
