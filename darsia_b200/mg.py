"""Device-resident V-cycle multigrid behind ``darsia_b200.MG`` (``utils/linear_solvers/mg.py:13-270``).

The cycle is host logic over C-ABI kernels that all run on the GPU: Jacobi sweeps
(``tvdk_jacobi_sweep``), the residual (``tvdk_residual``), restriction and prolongation
(``tvdk_mg_restrict`` / ``tvdk_mg_prolong_add``).  Every level is an ``Engine`` of its own (extents
halved, same batch of independent problems); nothing leaves the device during a solve unless a
tolerance asks for a norm.

Reference behaviour kept on purpose (see ``oracle.MG`` for the line-by-line restatement):
the ONE smoother object caches its diagonal at ``h = 1`` and is used on every level, so only the
residual operator sees the coarse spacing; the coarsest problem is smoothed from a zero guess; odd
extents lose their last entry in the restriction and get it back by edge padding; with ``tol`` the
previous iterate is returned on convergence.  Scalar coefficients only: the reference overwrites
per-voxel coefficients with block averages during its first cycle (``mg.py:134-149, 170-192``).
"""

from __future__ import annotations

import ctypes as C
from typing import List, Optional

import numpy as np
import torch

from ._runtime import device_guard
from .engine import Engine


class MGHierarchy:
    """Grids, buffers and the V-cycle for one batch of problems of a fixed shape."""

    def __init__(self, fine: Engine, depth: int, smoother_iterations: int, omega: float, ell: float) -> None:
        if depth < 0 or smoother_iterations < 1:
            raise ValueError("MG needs depth >= 0 and smoother_iterations >= 1")
        if np.ndim(omega) != 0 or np.ndim(ell) != 0:
            raise NotImplementedError(
                "MG: scalar mass / diffusion coefficients only (the reference replaces per-voxel coefficients by "
                "block averages during its first V-cycle, mg.py:134-149)"
            )
        self.depth, self.nsmooth = int(depth), int(smoother_iterations)
        self.omega, self.ell = float(omega), float(ell)
        self.levels: List[Engine] = [fine]
        for _ in range(self.depth + 1):  # depth + 1 restrictions: levels 0 .. depth + 1 (mg.py:215-223)
            prev = self.levels[-1]
            shape = tuple(n // 2 for n in prev.spatial)
            if min(shape) < 1:
                raise ValueError(
                    f"MG(depth={depth}) needs every extent >= {2 ** (self.depth + 1)}; got {fine.spatial} "
                    "(the reference fails on the empty coarse grid)"
                )
            self.levels.append(Engine(shape, nb=fine.nb, dtype=fine.np_dtype, device=fine.device))
        self.lib = fine.lib
        # per level: rhs (levels >= 1), two iterates for the sweep ping-pong, residual (all but the coarsest)
        self.rhs = [None] + [e.empty_fields(1) for e in self.levels[1:]]
        self.xa = [e.empty_fields(1) for e in self.levels]
        self.xb = [e.empty_fields(1) for e in self.levels]
        self.res = [e.empty_fields(1) for e in self.levels[:-1]]
        self.om = [e.coef(self.omega)[0] for e in self.levels]
        self.el = [e.coef(self.ell)[0] for e in self.levels]  # smoother: the h = 1 diagonal on every level
        self.el_h = [e.coef(self.ell / float(4 ** lv))[0] for lv, e in enumerate(self.levels)]  # residual: ell / h^2

    # -- building blocks ---------------------------------------------------------------------
    def _smooth(self, lv: int, x: torch.Tensor, rhs: torch.Tensor) -> torch.Tensor:
        """``smoother_iterations`` Jacobi sweeps starting from ``x`` (an external ``x`` is not modified);
        returns the buffer holding the result (``xa[lv]`` or ``xb[lv]``).  ``tvdk_jacobi_sweeps`` blocks
        the sweeps temporally (4 per pass in 2-D, 2 in 3-D)."""
        e, lib = self.levels[lv], self.lib
        ws = e.op_workspace()
        a, b = self.xa[lv], self.xb[lv]
        if x is b:
            a, b = b, a
        elif x is not a:
            a.copy_(x)
        in_b = C.c_int(0)
        with device_guard(e.device):
            rc = lib.jacobi_sweeps(C.byref(e.geom), C.byref(self.om[lv]), C.byref(self.el[lv]), a.data_ptr(),
                                   rhs.data_ptr(), b.data_ptr(), self.nsmooth, C.byref(in_b), ws.data_ptr(), ws.numel(),
                                   e.stream())
        lib.check(rc, "tvdk_jacobi_sweeps")
        return b if in_b.value else a

    def _vcycle(self, lv: int, x0: torch.Tensor, rhs: torch.Tensor, depth: int) -> torch.Tensor:
        """``base_V_Cycle`` (mg.py:194-243) on level ``lv``; returns the buffer holding the result."""
        e, lib = self.levels[lv], self.lib
        ec = self.levels[lv + 1]
        x = self._smooth(lv, x0, rhs)  # presmooth, mg.py:209
        ws = e.op_workspace()
        with device_guard(e.device):
            lib.check(lib.residual(C.byref(e.geom), C.byref(self.om[lv]), C.byref(self.el_h[lv]), x.data_ptr(),
                                   rhs.data_ptr(), self.res[lv].data_ptr(), ws.data_ptr(), ws.numel(), e.stream()),
                      "tvdk_residual")  # :212
            lib.check(lib.mg_restrict(C.byref(e.geom), C.byref(ec.geom), self.res[lv].data_ptr(),
                                      self.rhs[lv + 1].data_ptr(), e.stream()), "tvdk_mg_restrict")  # :215
        zero = self.xa[lv + 1]
        zero.zero_()
        if depth == 0:  # :218-219
            eps = self._smooth(lv + 1, zero, self.rhs[lv + 1])
        else:  # :220-223
            eps = self._vcycle(lv + 1, zero, self.rhs[lv + 1], depth - 1)
        with device_guard(e.device):
            lib.check(lib.mg_prolong_add(C.byref(e.geom), C.byref(ec.geom), eps.data_ptr(), x.data_ptr(), e.stream()),
                      "tvdk_mg_prolong_add")  # :226-239
        return self._smooth(lv, x, rhs)  # postsmooth, :242

    # -- the solver --------------------------------------------------------------------------
    def solve(self, x0: torch.Tensor, rhs: torch.Tensor, maxiter: int, tol: Optional[float]) -> torch.Tensor:
        """``MG.__call__`` (mg.py:245-270) on fields of the fine engine; returns a field tensor that may
        be ``x0`` itself (zero cycles) or one of the hierarchy's buffers (copy it before the next solve)."""
        e, lib = self.levels[0], self.lib
        x = x0
        if tol is None:
            for _ in range(int(maxiter)):
                x = self._vcycle(0, x, rhs, self.depth)
            return x
        if e.nb != 1:
            raise AssertionError("MG with tol works on one problem at a time")
        ws = e.op_workspace()
        dsum = torch.zeros(1, dtype=torch.float64, device=e.device)

        def norm2(a: torch.Tensor, b: Optional[torch.Tensor]) -> float:
            with device_guard(e.device):
                if b is None:
                    rc = lib.dot(C.byref(e.geom), a.data_ptr(), a.data_ptr(), dsum.data_ptr(), ws.data_ptr(), ws.numel(),
                                 e.stream())
                else:
                    rc = lib.norm_diff2(C.byref(e.geom), a.data_ptr(), b.data_ptr(), dsum.data_ptr(), ws.data_ptr(),
                                        ws.numel(), e.stream())
            lib.check(rc, "tvdk_dot")
            return float(dsum.item())

        x0n = float(np.sqrt(norm2(x0, None)))
        keep = e.empty_fields(1)  # the accepted iterate (the V-cycle's buffers are recycled)
        keep.copy_(x0)
        for _ in range(int(maxiter)):
            x_new = self._vcycle(0, keep, rhs, self.depth)
            with np.errstate(divide="ignore", invalid="ignore"):
                nrm = np.sqrt(norm2(x_new, keep)) / x0n
            if nrm < tol:  # mg.py:266-268: the previous iterate is returned
                break
            keep.copy_(x_new)
        return keep
