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
previous iterate is returned on convergence.

Per-voxel coefficients (``MG.heterogeneous``, decided by the constructor's arguments as in ``mg.py:70-73``):
the reference restricts the coefficient arrays on the way down and prolongates + edge-pads them on the way
up, overwriting its own attributes (``mg.py:134-149, 170-192``), and builds a fresh smoother at each of those
points (whose diagonal therefore DOES see the level's spacing).  Restricting a prolongated-and-padded array
gives the array back, so only two coefficient hierarchies ever occur: ``first`` -- the original arrays and
their pairwise averages, seen by the presmoothers and residuals of the very first cycle -- and ``stat`` -- the
block averages every later step sees.  Both are built once, on the device, with the same transfer kernels
the cycle uses; the residual multiplies a per-voxel ``ell`` after the Laplacian (``mg.py:104-106``).
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

    def __init__(self, fine: Engine, depth: int, smoother_iterations: int, omega, ell, heterogeneous: bool = False,
                 full_shape=None) -> None:
        if depth < 0 or smoother_iterations < 1:
            raise ValueError("MG needs depth >= 0 and smoother_iterations >= 1")
        self.hetero = bool(heterogeneous)
        if not self.hetero and (np.ndim(omega) != 0 or np.ndim(ell) != 0):
            # what the reference does: ``heterogeneous`` is False, nothing restricts the arrays, and the first coarse
            # smoother call dies in NumPy's broadcasting (fine-grid coefficients against a coarse right-hand side)
            raise ValueError(
                "MG was constructed without per-voxel coefficients (heterogeneous = False, mg.py:70-73) but the solve has "
                "per-voxel omega / ell: operands could not be broadcast together on the coarse grid.  Construct it as the "
                "reference requires, MG(mass_coeff=omega, diffusion_coeff=ell, ...)"
            )
        self.depth, self.nsmooth = int(depth), int(smoother_iterations)
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
        self.cycles = 0  # V-cycles run so far: the first one descends through the `first` hierarchy
        self._keep = []  # scaled coefficient fields (the tvdk_coef_t structs hold raw pointers)
        self.maps = {}  # name -> (block-averaged fine-level field, problems it spans)
        nlev = len(self.levels)
        if not self.hetero:
            self.omega, self.ell = float(omega), float(ell)
            om = [e.coef(self.omega)[0] for e in self.levels]
            el = [e.coef(self.ell)[0] for e in self.levels]  # smoother: the h = 1 diagonal on every level
            el_h = [e.coef(self.ell / float(4 ** lv))[0] for lv, e in enumerate(self.levels)]  # residual: ell / h^2
            self.coefs = {w: [(om[lv], el[lv], el_h[lv]) for lv in range(nlev)] for w in ("first", "stat")}
            return
        om_f, om_s = self._chains(omega, full_shape, "omega", scale=False)
        el_f, el_s = self._chains(ell, full_shape, "ell", scale=True)
        # (a fresh smoother per level: its diagonal is cached at that level's h, so smoother and residual share ell / h^2)
        self.coefs = {"first": [(om_f[lv], el_f[lv], el_f[lv]) for lv in range(nlev)],
                      "stat": [(om_s[lv], el_s[lv], el_s[lv]) for lv in range(nlev)]}

    def _chains(self, value, full_shape, name: str, scale: bool):
        """The two hierarchies of one coefficient: ``tvdk_coef_t`` per level for ``first`` (the value and its successive
        restrictions, mg.py:134-142) and ``stat`` (the coarsest of those prolongated + edge-padded back up, :170-185).
        ``scale``: the coefficient enters as ``ell / h^2`` (h = 2^level; a power of four, exact)."""
        fine, lib = self.levels[0], self.lib
        nlev = len(self.levels)
        if np.ndim(value) == 0 and not isinstance(value, torch.Tensor):
            cs = [e.coef(float(value) / (float(4 ** lv) if scale else 1.0))[0] for lv, e in enumerate(self.levels)]
            return cs, cs
        c0, field0 = fine.coef(value, full_shape, name)
        if field0 is None:  # a 0-d tensor
            return self._chains(float(c0.value), full_shape, name, scale)
        nbc = fine.nb if c0.batch_stride else 1
        geoms = [e._geom_nb(nbc) for e in self.levels]
        first = [field0]
        for lv in range(nlev - 1):
            e = self.levels[lv]
            coarse = torch.zeros(self.levels[lv + 1].field_elems * nbc, dtype=fine.tdtype, device=fine.device)
            with device_guard(e.device):
                lib.check(lib.mg_restrict(C.byref(geoms[lv]), C.byref(geoms[lv + 1]), first[lv].data_ptr(), coarse.data_ptr(),
                                          e.stream()), "tvdk_mg_restrict")
            first.append(coarse)
        stat = [None] * nlev
        stat[-1] = first[-1]
        for lv in range(nlev - 2, -1, -1):
            e = self.levels[lv]
            up = torch.zeros(e.field_elems * nbc, dtype=fine.tdtype, device=fine.device)
            with device_guard(e.device):
                lib.check(lib.mg_prolong_add(C.byref(geoms[lv]), C.byref(geoms[lv + 1]), stat[lv + 1].data_ptr(), up.data_ptr(),
                                             e.stream()), "tvdk_mg_prolong_add")
            stat[lv] = up
        self.maps[name] = (stat[0], nbc)  # what the reference leaves in mass_coeff / diffusion_coeff after a cycle

        def as_coefs(fields):
            out = []
            for lv, (e, fld) in enumerate(zip(self.levels, fields)):
                if scale and lv > 0:
                    fld = fld * (1.0 / float(4 ** lv))  # (set-up only; exact)
                self._keep.append(fld)
                out.append(e.coef(fld)[0])
            return out

        return as_coefs(first), as_coefs(stat)

    def final_coefficient(self, name: str, shape) -> np.ndarray:
        """Host copy of the block-averaged map the reference's solver object holds after its first cycle."""
        fld, nbc = self.maps[name]
        e = self.levels[0]
        inter = e.unpack(fld, nb=nbc)
        return inter.cpu().numpy().reshape(shape)

    # -- building blocks ---------------------------------------------------------------------
    def _smooth(self, lv: int, x: torch.Tensor, rhs: torch.Tensor, which: str) -> torch.Tensor:
        """``smoother_iterations`` Jacobi sweeps starting from ``x`` (an external ``x`` is not modified);
        returns the buffer holding the result (``xa[lv]`` or ``xb[lv]``).  ``tvdk_jacobi_sweeps`` blocks
        the sweeps temporally (4 per pass in 2-D, 2 in 3-D) when the coefficients are scalars."""
        e, lib = self.levels[lv], self.lib
        om, el, _ = self.coefs[which][lv]
        ws = e.op_workspace()
        a, b = self.xa[lv], self.xb[lv]
        if x is b:
            a, b = b, a
        elif x is not a:
            a.copy_(x)
        in_b = C.c_int(0)
        with device_guard(e.device):
            rc = lib.jacobi_sweeps(C.byref(e.geom), C.byref(om), C.byref(el), a.data_ptr(),
                                   rhs.data_ptr(), b.data_ptr(), self.nsmooth, C.byref(in_b), ws.data_ptr(), ws.numel(),
                                   e.stream())
        lib.check(rc, "tvdk_jacobi_sweeps")
        return b if in_b.value else a

    def _vcycle(self, lv: int, x0: torch.Tensor, rhs: torch.Tensor, depth: int, down: str) -> torch.Tensor:
        """``base_V_Cycle`` (mg.py:194-243) on level ``lv``; returns the buffer holding the result.  ``down``: the
        coefficient hierarchy the presmoother and the residual see (``first`` during the very first cycle)."""
        e, lib = self.levels[lv], self.lib
        ec = self.levels[lv + 1]
        x = self._smooth(lv, x0, rhs, down)  # presmooth, mg.py:209
        om, _, el_h = self.coefs[down][lv]
        ws = e.op_workspace()
        with device_guard(e.device):
            lib.check(lib.residual(C.byref(e.geom), C.byref(om), C.byref(el_h), x.data_ptr(),
                                   rhs.data_ptr(), self.res[lv].data_ptr(), ws.data_ptr(), ws.numel(), e.stream()),
                      "tvdk_residual")  # :212
            lib.check(lib.mg_restrict(C.byref(e.geom), C.byref(ec.geom), self.res[lv].data_ptr(),
                                      self.rhs[lv + 1].data_ptr(), e.stream()), "tvdk_mg_restrict")  # :215
        zero = self.xa[lv + 1]
        zero.zero_()
        if depth == 0:  # :218-219 (the coarsest level's coefficients are the same in both hierarchies)
            eps = self._smooth(lv + 1, zero, self.rhs[lv + 1], down)
        else:  # :220-223
            eps = self._vcycle(lv + 1, zero, self.rhs[lv + 1], depth - 1, down)
        with device_guard(e.device):
            lib.check(lib.mg_prolong_add(C.byref(e.geom), C.byref(ec.geom), eps.data_ptr(), x.data_ptr(), e.stream()),
                      "tvdk_mg_prolong_add")  # :226-239
        return self._smooth(lv, x, rhs, "stat")  # postsmooth, :242 (after prolongate_parameters)

    def _cycle(self, x: torch.Tensor, rhs: torch.Tensor) -> torch.Tensor:
        out = self._vcycle(0, x, rhs, self.depth, "first" if self.cycles == 0 else "stat")
        self.cycles += 1
        return out

    # -- the solver --------------------------------------------------------------------------
    def solve(self, x0: torch.Tensor, rhs: torch.Tensor, maxiter: int, tol: Optional[float]) -> torch.Tensor:
        """``MG.__call__`` (mg.py:245-270) on fields of the fine engine; returns a field tensor that may
        be ``x0`` itself (zero cycles) or one of the hierarchy's buffers (copy it before the next solve)."""
        e, lib = self.levels[0], self.lib
        x = x0
        if tol is None:
            for _ in range(int(maxiter)):
                x = self._cycle(x, rhs)
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
            x_new = self._cycle(keep, rhs)
            with np.errstate(divide="ignore", invalid="ignore"):
                nrm = np.sqrt(norm2(x_new, keep)) / x0n
            if nrm < tol:  # mg.py:266-268: the previous iterate is returned
                break
            keep.copy_(x_new)
        return keep
