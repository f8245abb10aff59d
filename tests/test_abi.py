"""CPU: the C-ABI library loads and exports every symbol include/b200mf.h
declares; the Python binding table covers the same set (no compute calls)."""
import ctypes
import os
import re

from b200mf import lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "b200mf.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(b200mf_[a-zA-Z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    assert os.path.exists(L.LIB_PATH), "libb200mf.so not built (run __graft_entry__.build())"
    dll = ctypes.CDLL(L.LIB_PATH)
    syms = header_symbols()
    assert len(syms) > 30
    missing = [s for s in syms if not hasattr(dll, s)]
    assert not missing, missing


def test_python_binding_covers_header():
    assert sorted(L.SIGNATURES) == header_symbols()
    lib = L.load()
    assert lib.b200mf_version() >= 100
    assert lib.b200mf_launch_count() == 0 or lib.b200mf_launch_count() > 0


def test_unknown_caching_is_rejected_on_host():
    import pytest
    from b200mf import NeoHookOperator

    op = NeoHookOperator()
    with pytest.raises(ValueError, match="Unknown caching"):
        op.initialize(None, "tensor3")
