"""CPU-side checks of the drop-in boundary: the shared library loads and exports every symbol
include/hsbp.h declares, the ctypes table covers all of them, and -- without a GPU -- context
creation fails loudly instead of falling back to anything."""
import ctypes
import os

import pytest

import hybridsbp_b200 as hs


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as g
    g.build()
    return hs.lib()


def test_library_exports_every_declared_symbol(built):
    names = hs.declared_symbols()
    assert len(names) >= 25
    raw = ctypes.CDLL(hs.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), "libhsbp.so does not export %s" % n


def test_ctypes_table_matches_header(built):
    assert sorted(built._signatures) == hs.declared_symbols()


def test_no_cpu_fallback(built):
    from tests.conftest import _has_gpu
    if _has_gpu():
        pytest.skip("GPU present")
    with pytest.raises(hs.HsbpError):
        hs.Context(0)


def test_product_does_not_import_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for dp, _, files in os.walk(os.path.join(root, "hybridsbp_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f
