"""CPU-only: libedelweiss_b200.so loads and exports every symbol include/edelweiss_b200.h declares."""
import ctypes
import os
import re

from conftest import ROOT

from edelweissfe_b200 import _lib


def _declared():
    src = open(os.path.join(ROOT, "include", "edelweiss_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ewb_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    names = _declared()
    assert len(names) >= 15
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), n


def test_binding_covers_header():
    assert sorted(_lib.SYMBOLS) == _declared()
    lib = _lib.load()
    assert lib.ewb_version() >= 100
    assert lib.ewb_launch_count() >= 0
