"""The C-ABI library loads on a CPU-only box and exports every function include/nfk.h declares."""
import ctypes
import os
import re

import pytest

from conftest import ROOT
from nflows_b200 import _native


def declared_functions():
    text = open(os.path.join(ROOT, "include", "nfk.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(nfk_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    assert declared_functions() == sorted(_native.EXPORTED_SYMBOLS)


def test_library_exports_every_symbol():
    if not os.path.exists(_native.library_path()):
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(_native.library_path())
    for name in declared_functions():
        assert hasattr(lib, name), name
    loaded = _native.load()
    assert loaded.nfk_version() == 2
    assert loaded.nfk_launch_count() >= 0


def test_no_silent_fallback_on_missing_library(monkeypatch):
    monkeypatch.setattr(_native, "_lib", None)
    monkeypatch.setattr(_native, "_LIB_PATH", "/nonexistent/libnfk_sm100.so")
    with pytest.raises(_native.NativeUnavailable):
        _native.load()
