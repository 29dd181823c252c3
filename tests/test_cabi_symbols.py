"""CPU-side checks of the drop-in boundary: the C-ABI library builds for sm_100a (nvcc
cross-compiles without a GPU), loads, and exports every symbol include/linkteller_b200.h declares;
the ctypes stub binds exactly that set.  No compute calls are made here."""
import ctypes
import os

import pytest

from linkteller_b200 import _lib
from linkteller_b200.build import build_library


@pytest.fixture(scope="module")
def lib_path():
    return build_library()


def test_header_symbols_all_exported(lib_path):
    lib = ctypes.CDLL(lib_path)
    names = _lib.header_symbols()
    assert len(names) >= 25
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in the header but not exported: {missing}"


def test_ctypes_stub_covers_header(lib_path):
    assert sorted(_lib._SIGS) == _lib.header_symbols()
    lib = _lib.load()
    assert lib.lt_version() >= 100
    assert lib.lt_last_error() is not None


def test_built_for_sm_100a(lib_path):
    log = os.path.join(os.path.dirname(lib_path), "build.log")
    assert "arch=compute_100a,code=sm_100a" in open(log).read()


def test_product_never_imports_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "linkteller_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith(".py"):
                src = open(os.path.join(dp, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f
