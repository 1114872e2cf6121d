"""CPU: the sm_100a shared library loads and exports every symbol include/tvdk.h declares
(no compute calls -- there is no GPU here), and the product refuses to run without a GPU."""

import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT
from darsia_b200 import _cabi

LIB = _cabi.LIB_PATH
needs_lib = pytest.mark.skipif(not os.path.exists(LIB), reason="libtvdk.so not built (run __graft_entry__.build())")


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "tvdk.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tvdk_[a-z0-9_]+)\s*\(", text)))


def test_binding_covers_header():
    assert declared_symbols() == sorted(_cabi.SIGNATURES)


@needs_lib
def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(LIB)
    for name in declared_symbols():
        assert hasattr(lib, name), name
    assert not hasattr(lib, "tvdk_is_emulator")  # the shipped library is not the CPU stepper
    wrapped = _cabi.Library(LIB)
    assert wrapped.version() == 100
    assert wrapped.pitch(_cabi.TVDK_F64, 37) == 48 and wrapped.pitch(_cabi.TVDK_F32, 37) == 64
    assert not wrapped.is_emulator


def test_struct_layouts_match_header():
    # sizes the C side computes with sizeof(tvdk_stat_t)/sizeof(double|int)
    assert ctypes.sizeof(_cabi.Stat) == 12 * 8 + 8 * 4 + 2 * 8
    assert ctypes.sizeof(_cabi.Geom) == 11 * 4
    assert ctypes.sizeof(_cabi.Coef) == 24


@needs_lib
def test_no_cpu_fallback(monkeypatch):
    """Without a CUDA device the product path raises; it never computes on the host."""
    import torch

    import darsia_b200 as da

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    monkeypatch.setattr(_cabi, "_LIB", None)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        da.split_bregman_tvd(np.zeros((8, 8)))


def test_missing_library_fails_loudly(tmp_path):
    with pytest.raises(ImportError, match="no CPU fallback"):
        _cabi.Library(str(tmp_path / "libtvdk.so"))
