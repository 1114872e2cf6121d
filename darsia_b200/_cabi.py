"""ctypes binding of ``libtvdk.so`` (the C ABI declared in ``include/tvdk.h``).

This is the binding a DarSIA maintainer would add on the reference side:
plain pointers and sizes, no torch types in any signature.  The product path has NO CPU
fallback: if the CUDA library is missing, or the current device is not a B200-class
(sm_100) GPU, loading fails loudly.
"""

from __future__ import annotations

import ctypes as C
import os
from typing import Optional

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "csrc", "libtvdk.so")
# Tuning runs (tools/ab.sh) compare alternative BUILDS of the same CUDA sources: TVDK_LIB may name another
# libtvdk*.so inside darsia_b200/csrc/ -- nothing else, and there is still no CPU path.
_alt = os.environ.get("TVDK_LIB")
if _alt:
    _alt = os.path.abspath(_alt)
    if os.path.dirname(_alt) == os.path.join(HERE, "csrc") and os.path.basename(_alt).startswith("libtvdk"):
        LIB_PATH = _alt
    else:
        raise ImportError(f"TVDK_LIB={_alt!r}: only alternative builds libtvdk*.so inside darsia_b200/csrc are accepted")

TVDK_F64, TVDK_F32 = 0, 1
TVDK_SOLVER_JACOBI, TVDK_SOLVER_CG = 0, 1


class Geom(C.Structure):
    """``tvdk_geom_t``."""

    _fields_ = [
        ("dim", C.c_int32),
        ("dtype", C.c_int32),
        ("n", C.c_int32 * 3),
        ("pitch", C.c_int32),
        ("nb", C.c_int32),
        ("g0_off", C.c_int32),
        ("g0_n", C.c_int32),
        ("z_begin", C.c_int32),
        ("z_end", C.c_int32),
    ]


class Coef(C.Structure):
    """``tvdk_coef_t``."""

    _fields_ = [("device_ptr", C.c_void_p), ("value", C.c_double), ("batch_stride", C.c_int64)]


class Stat(C.Structure):
    """``tvdk_stat_t``."""

    _fields_ = [
        ("nrm_f", C.c_double),
        ("inc", C.c_double),
        ("energy", C.c_double),
        ("bnrm", C.c_double),
        ("atol", C.c_double),
        ("rho", C.c_double),
        ("rho_prev", C.c_double),
        ("pq", C.c_double),
        ("alpha", C.c_double),
        ("beta", C.c_double),
        ("jac_x0nrm2", C.c_double),
        ("jac_err", C.c_double),
        ("iters", C.c_int32),
        ("done", C.c_int32),
        ("cg_stop", C.c_int32),
        ("cg_first", C.c_int32),
        ("cg_iters", C.c_int32),
        ("jac_stop", C.c_int32),
        ("jac_sel", C.c_int32),
        ("cur_buf", C.c_int32),
        ("e_init", C.c_double),
        ("e_prev", C.c_double),
        ("rho_pred", C.c_double),
        ("cg_pred_dev", C.c_double),
        ("nonfinite", C.c_int32),
        ("pad_", C.c_int32),
    ]


class Problem(C.Structure):
    """``tvdk_problem_t``."""

    _fields_ = [
        ("geom", Geom),
        ("isotropic", C.c_int32),
        ("solver_kind", C.c_int32),
        ("solver_maxiter", C.c_int32),
        ("max_num_iter", C.c_int32),
        ("solver_tol", C.c_double),
        ("eps", C.c_double),
        ("mu", Coef),
        ("omega", Coef),
        ("ell", Coef),
        ("f", C.c_void_p),
        ("u", C.c_void_p),
        ("t", C.c_void_p),
        ("t_comp_stride", C.c_int64),
        ("x0_d", C.c_void_p),
        ("x0_b", C.c_void_p),
        ("x0_comp_stride", C.c_int64),
        ("workspace", C.c_void_p),
        ("workspace_bytes", C.c_size_t),
        ("stats", C.c_void_p),
        ("history", C.c_void_p),
        ("energy_history", C.c_void_p),
        ("verbose_every", C.c_int32),
        ("check_every", C.c_int32),
        ("fused", C.c_int32),
        ("tile_x", C.c_int32),
        ("tile_y", C.c_int32),
        ("zchunk", C.c_int32),
        ("adaptive_mask", C.c_void_p),
    ]


class Chambolle(C.Structure):
    """``tvdk_chambolle_t``."""

    _fields_ = [
        ("geom", Geom),
        ("weight", C.c_double),
        ("eps", C.c_double),
        ("max_num_iter", C.c_int32),
        ("check_every", C.c_int32),
        ("f", C.c_void_p),
        ("out", C.c_void_p),
        ("p0", C.c_void_p),
        ("p1", C.c_void_p),
        ("p_comp_stride", C.c_int64),
        ("workspace", C.c_void_p),
        ("workspace_bytes", C.c_size_t),
        ("stats", C.c_void_p),
        ("history", C.c_void_p),
    ]


TVDK_MAX_SLABS, TVDK_SLAB_GL, TVDK_SLAB_GH = 16, 1, 2


class SlabComm(C.Structure):
    """``tvdk_slab_comm_t``."""

    _fields_ = [
        ("world", C.c_int32),
        ("rank", C.c_int32),
        ("nown", C.c_int32 * TVDK_MAX_SLABS),
        ("arena", C.c_void_p * TVDK_MAX_SLABS),
        ("arena_bytes", C.c_size_t),
        ("seq", C.POINTER(C.c_uint64)),
        ("timeout_ms", C.c_int32),
    ]


_PG = C.POINTER(Geom)
_PC = C.POINTER(Coef)
_VP = C.c_void_p
_SZ = C.c_size_t
_I64 = C.c_int64
_I = C.c_int

# every symbol include/tvdk.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "tvdk_version": (C.c_int, []),
    "tvdk_last_error": (C.c_char_p, []),
    "tvdk_device_ok": (C.c_int, []),
    "tvdk_launch_count": (_I64, [_I]),
    "tvdk_set_tune": (C.c_int, [_I, _I, _I]),
    "tvdk_pitch": (C.c_int, [_I, _I]),
    "tvdk_field_elems": (_I64, [_PG]),
    "tvdk_workspace_bytes": (_SZ, [C.POINTER(Problem)]),
    "tvdk_op_workspace_bytes": (_SZ, [_PG]),
    "tvdk_solve": (C.c_int, [C.POINTER(Problem), _VP, C.POINTER(C.c_int)]),
    "tvdk_pack": (C.c_int, [_PG, _I, _VP, _VP, _VP]),
    "tvdk_unpack": (C.c_int, [_PG, _I, _VP, _VP, _VP]),
    "tvdk_rhs": (C.c_int, [_PG, _I, _PC, _PC, _PC, _VP, _VP, _I64, _VP, _VP, _SZ, _VP]),
    "tvdk_jacobi_sweep": (C.c_int, [_PG, _PC, _PC, _VP, _VP, _VP, _VP, _SZ, _VP]),
    "tvdk_cg_matvec": (C.c_int, [_PG, _PC, _PC, _VP, _VP, _VP, _SZ, _VP]),
    "tvdk_shrink": (C.c_int, [_PG, _I, _PC, _PC, _VP, _VP, _VP, _VP, _VP, _I64, _VP, _VP, _SZ, _VP]),
    "tvdk_rhs_explicit": (C.c_int, [_PG, _PC, _PC, _VP, _VP, _VP, _I64, _VP, _VP, _SZ, _VP]),
    "tvdk_shrink_explicit": (C.c_int, [_PG, _I, _PC, _PC, _VP, _VP, _VP, _VP, _I64, _VP, _I64, _VP, _VP, _SZ, _VP]),
    "tvdk_expand_state": (C.c_int, [_PG, _I, _PC, _PC, _VP, _I64, _VP, _VP, _I64, _VP]),
    "tvdk_dot": (C.c_int, [_PG, _VP, _VP, _VP, _VP, _SZ, _VP]),
    "tvdk_norm_diff2": (C.c_int, [_PG, _VP, _VP, _VP, _VP, _SZ, _VP]),
    "tvdk_energy": (C.c_int, [_PG, _PC, _PC, _VP, _VP, _VP, _VP, _SZ, _VP]),
    "tvdk_fused_jacobi1": (C.c_int, [_PG, _I, _PC, _PC, _PC, _VP, _VP, _VP, _VP, _VP, _I64, _VP, _VP, _SZ, _VP]),
    "tvdk_fused_jacobi1_x2": (C.c_int, [_PG, _I, _PC, _PC, _PC, _VP, _VP, _VP, _VP, _VP, _I64, _VP, _VP, _SZ, _VP]),
    "tvdk_cg_init": (C.c_int, [_PG, _I, _PC, _PC, _PC, _VP, _VP, _I64, _VP, _VP, _VP, _VP, _VP, _SZ, _VP]),
    "tvdk_cg_pq": (C.c_int, [_PG, _PC, _PC, _VP, _VP, _VP, _VP, _I, _VP, _VP, _SZ, _VP]),
    "tvdk_cg_update": (C.c_int, [_PG, _PC, _PC, _VP, _VP, _VP, _VP, _VP, _VP, _SZ, _VP]),
    "tvdk_jacobi_sweeps": (C.c_int, [_PG, _PC, _PC, _VP, _VP, _VP, _I, C.POINTER(C.c_int), _VP, _SZ, _VP]),
    "tvdk_chambolle": (C.c_int, [C.POINTER(Chambolle), _VP, C.POINTER(C.c_int)]),
    "tvdk_residual": (C.c_int, [_PG, _PC, _PC, _VP, _VP, _VP, _VP, _SZ, _VP]),
    "tvdk_mg_restrict": (C.c_int, [_PG, _PG, _VP, _VP, _VP]),
    "tvdk_mg_prolong_add": (C.c_int, [_PG, _PG, _VP, _VP, _VP]),
    "tvdk_cg_first": (C.c_int, [_PG, _PC, _PC, _VP, _VP, _VP, _SZ, _VP]),
    "tvdk_cg_fused": (C.c_int, [_PG, _PC, _PC, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _SZ, _VP]),
    "tvdk_slab_header_bytes": (_SZ, []),
    "tvdk_slab_workspace_bytes": (_SZ, [C.POINTER(Problem)]),
    "tvdk_solve_slab": (C.c_int, [_I, C.POINTER(C.POINTER(Problem)), C.POINTER(C.POINTER(SlabComm)), C.POINTER(_VP),
                                  C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "tvdk_slab_error": (C.c_int, [_VP, _VP, C.POINTER(_I64)]),
    "tvdk_slab_exchange": (C.c_int, [_I, C.POINTER(_PG), C.POINTER(C.POINTER(SlabComm)), _I, C.POINTER(C.POINTER(_VP)),
                                     _I, _I, C.POINTER(_VP), C.POINTER(C.c_int)]),
    "tvdk_set_pdl": (C.c_int, [_I]),
    "tvdk_profile_arm": (C.c_int, [C.c_char_p]),
    "tvdk_profile_read": (C.c_int, [C.POINTER(_I64), C.POINTER(C.c_double)]),
    "tvdk_ipc_export": (C.c_int, [_VP, C.POINTER(C.c_ubyte * 64), C.POINTER(_I64)]),
    "tvdk_ipc_open": (C.c_int, [C.POINTER(C.c_ubyte * 64), _I64, C.POINTER(_VP)]),
    "tvdk_ipc_close": (C.c_int, [_VP, _I64]),
    "tvdk_enable_peer": (C.c_int, [_I]),
}


class TvdkError(RuntimeError):
    """A libtvdk call returned a non-zero code."""


class Library:
    """A loaded libtvdk with typed entry points; ``check`` raises on non-zero return codes."""

    def __init__(self, path: str) -> None:
        if not os.path.exists(path):
            raise ImportError(
                f"{path} not found: build the CUDA extension first "
                "(python -c 'import __graft_entry__ as g; g.build()' or make -C darsia_b200/csrc). "
                "darsia_b200 has no CPU fallback."
            )
        self.path = path
        self.lib = C.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(self.lib, name)  # AttributeError if the symbol is missing
            fn.restype = res
            fn.argtypes = args
            setattr(self, name[5:], fn)
        self.is_emulator = hasattr(self.lib, "tvdk_is_emulator")

    def check(self, rc: int, what: str) -> None:
        if rc != 0:
            msg = self.last_error().decode("utf-8", "replace")
            kind = "CUDA error" if rc > 0 else "tvdk error"
            raise TvdkError(f"{what}: {kind} {rc}: {msg}")


_LIB: Optional[Library] = None


def get_library() -> Library:
    """The process-wide library handle (loaded on first use; raises if it cannot be)."""
    global _LIB
    if _LIB is None:
        _LIB = Library(LIB_PATH)
    return _LIB
