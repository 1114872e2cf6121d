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
    assert ctypes.sizeof(_cabi.Stat) == 12 * 8 + 8 * 4 + 4 * 8 + 2 * 4  # (+ nonfinite, pad)
    assert ctypes.sizeof(_cabi.SlabComm) == 8 + 16 * 4 + 16 * 8 + 8 + 8 + 8
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


@needs_lib
def test_argument_validation_needs_no_gpu():
    """Every entry point validates its arguments before touching the device: error codes and messages of the
    real sm_100a library, checked here on the CPU (no kernel is launched by any of these calls)."""
    lib = _cabi.Library(LIB)
    ERR_INVALID, ERR_UNSUPPORTED, ERR_WORKSPACE = -1, -2, -3

    def geom(dim=2, n=(16, 24, 1), pitch=None, nb=1, dtype=_cabi.TVDK_F64):
        g = _cabi.Geom()
        g.dim, g.dtype, g.nb = dim, dtype, nb
        for a in range(3):
            g.n[a] = n[a]
        g.pitch = lib.pitch(dtype, n[min(dim, 3) - 1]) if pitch is None else pitch
        g.g0_off, g.g0_n, g.z_begin, g.z_end = 0, n[0], 0, n[0]
        return g

    assert lib.solve(None, None, None) == ERR_INVALID and b"NULL" in lib.last_error()
    p = _cabi.Problem()
    p.geom = geom(dim=4)
    assert lib.solve(ctypes.byref(p), None, None) == ERR_INVALID and b"dim must be 2 or 3" in lib.last_error()
    p.geom = geom(pitch=25)  # not a multiple of 16 bytes
    assert lib.solve(ctypes.byref(p), None, None) == ERR_INVALID and b"pitch" in lib.last_error()
    p.geom = geom(nb=70000)
    assert lib.solve(ctypes.byref(p), None, None) == ERR_INVALID and b"nb" in lib.last_error()
    p.geom = geom()
    p.max_num_iter, p.solver_kind = 3, 7
    assert lib.solve(ctypes.byref(p), None, None) == ERR_UNSUPPORTED and b"solver kind" in lib.last_error()
    p.solver_kind, p.solver_maxiter, p.solver_tol = _cabi.TVDK_SOLVER_CG, 3, 0.0
    p.mu.value = p.omega.value = p.ell.value = 1.0
    need = lib.workspace_bytes(ctypes.byref(p))  # host arithmetic only
    field = 16 * lib.pitch(_cabi.TVDK_F64, 24) * 8
    assert 4 * field < need < 5 * field  # X, R, P0, P1 + scratch (two-kernel CG route: p.fused == 0)
    p.fused = 1
    need = lib.workspace_bytes(ctypes.byref(p))
    assert 5 * field < need < 6 * field  # + R1: the single-pass route ping-pongs r as well
    p.workspace, p.workspace_bytes = 1, need - 1  # (the pointer is never dereferenced: the size check comes first)
    assert lib.solve(ctypes.byref(p), None, None) == ERR_WORKSPACE and b"workspace" in lib.last_error()
    p.workspace_bytes = need
    assert lib.solve(ctypes.byref(p), None, None) == ERR_INVALID and b"non-NULL" in lib.last_error()  # f, u, t, stats
    # operators
    g = geom()
    assert lib.fused_jacobi1(ctypes.byref(g), 0, None, None, None, 8, 16, 16, 24, 32, 0, None, None, 0, None) == ERR_INVALID
    assert b"out-of-place" in lib.last_error()
    assert lib.rhs_explicit(ctypes.byref(g), None, None, 8, None, 16, 0, 24, None, 0, None) == ERR_INVALID
    assert b"both d and b" in lib.last_error()
    assert lib.shrink_explicit(ctypes.byref(g), 0, None, None, 8, 16, 16, None, 0, 24, 0, None, None, 0, None) == ERR_INVALID
    assert b"b_in" in lib.last_error()
    assert lib.ipc_close(None, 0) == ERR_INVALID
    gc = geom(n=(7, 12, 1))  # not fine / 2
    assert lib.mg_restrict(ctypes.byref(g), ctypes.byref(gc), 8, 16, None) == ERR_INVALID and b"coarse extent" in lib.last_error()
    c = _cabi.Chambolle()
    c.geom, c.max_num_iter, c.weight = geom(), 0, 0.1
    assert lib.chambolle(ctypes.byref(c), None, None) == ERR_INVALID and b"max_num_iter" in lib.last_error()
    assert lib.field_elems(ctypes.byref(geom(dim=3, n=(4, 5, 6)))) == 4 * 5 * lib.pitch(_cabi.TVDK_F64, 6)
