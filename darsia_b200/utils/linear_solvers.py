"""``darsia.Solver`` / ``darsia.Jacobi`` / ``darsia.CG`` / ``darsia.MG`` with the reference's signatures.

Reference: ``utils/linear_solvers/solver.py:12-72``, ``jacobi.py:14-133``, ``cg.py:10-59``, ``mg.py:13-270``.
Inside ``split_bregman_tvd`` the objects are *descriptors* (kind, ``maxiter``, ``tol``): the
inner solves run inside ``libtvdk``'s kernels, never through ``__call__``.  ``__call__`` is
still provided -- it solves ``mass_coeff * x - diffusion_coeff * laplace(x) = rhs`` on the GPU
with the same kernels -- so that code using the solvers standalone (e.g. DarSIA's
``H1_regularization``, ``restoration/h1_regularization.py:54-67``) keeps working.

Differences from the reference, all from SURVEY.md's quirk table: no state is cached between
calls (Q2); ``CG.tol`` is a relative tolerance and must be a number (Q4); arrays with more
axes than ``dim`` are treated as independent problems (Q5/Q6).
"""

from __future__ import annotations

import ctypes as C
from typing import Optional, Union

import numpy as np

from .. import _cabi

Coeff = Union[float, int, np.ndarray]


class Solver:
    """Parameter carrier for ``mass_coeff * x - diffusion_coeff * laplace(x) = rhs``."""

    kind: Optional[int] = None

    def __init__(
        self,
        maxiter: int = 1,
        tol: Optional[float] = None,
        dim: int = 2,
        mass_coeff: Optional[Coeff] = None,
        diffusion_coeff: Optional[Coeff] = None,
        verbose=False,
    ) -> None:
        self.dim = dim
        self.maxiter = maxiter
        self.tol = tol
        self.mass_coeff = mass_coeff
        self.diffusion_coeff = diffusion_coeff
        self.verbose = verbose

    def update_params(
        self,
        dim: Optional[int] = None,
        mass_coeff: Optional[Coeff] = None,
        diffusion_coeff: Optional[Coeff] = None,
    ) -> None:
        self.dim = dim if dim is not None else self.dim
        self.mass_coeff = mass_coeff if mass_coeff is not None else self.mass_coeff
        self.diffusion_coeff = diffusion_coeff if diffusion_coeff is not None else self.diffusion_coeff

    def __call__(self, x0: np.ndarray, rhs: np.ndarray) -> np.ndarray:
        raise NotImplementedError(
            f"{type(self).__name__} has no GPU implementation (darsia_b200 implements Jacobi and CG and has no "
            "CPU fallback)"
        )

    # -- shared by Jacobi / CG ---------------------------------------------------------------
    def _engine(self, x0: np.ndarray, rhs: np.ndarray):
        from ..engine import Engine

        assert x0.shape == rhs.shape
        if not np.issubdtype(x0.dtype, np.floating):
            raise TypeError("float arrays only")
        dtype = np.float32 if x0.dtype == np.float32 else np.float64
        spatial = x0.shape[: self.dim]
        nb = int(np.prod(x0.shape[self.dim :], dtype=np.int64))
        eng = Engine(spatial, nb=nb, dtype=dtype)
        om, k1 = eng.coef(self.mass_coeff, x0.shape, "mass_coeff")
        el, k2 = eng.coef(self.diffusion_coeff, x0.shape, "diffusion_coeff")
        return eng, om, el, (k1, k2)


class Jacobi(Solver):
    """Matrix-free Jacobi sweeps (``jacobi.py:74-133``)."""

    kind = _cabi.TVDK_SOLVER_JACOBI

    def __call__(self, x0: np.ndarray, rhs: np.ndarray, h: float = 1.0) -> np.ndarray:
        if h != 1.0:
            raise NotImplementedError("only h = 1.0 (the value the TVD path uses) is implemented")
        import torch

        eng, om, el, _keep = self._engine(x0, rhs)
        lib = eng.lib
        x = eng.upload(x0.astype(eng.np_dtype, copy=False))
        b = eng.upload(rhs.astype(eng.np_dtype, copy=False))
        y = eng.empty_fields(1)
        ws = eng.op_workspace()
        x0n = None
        if self.tol is not None:
            x0n = float(np.linalg.norm(x0))
        if self.tol is None:  # fixed count: one call, temporally blocked sweeps
            in_b = C.c_int(0)
            rc = lib.jacobi_sweeps(C.byref(eng.geom), C.byref(om), C.byref(el), x.data_ptr(), b.data_ptr(), y.data_ptr(),
                                   int(self.maxiter), C.byref(in_b), ws.data_ptr(), ws.numel(), eng.stream())
            lib.check(rc, "tvdk_jacobi_sweeps")
            return eng.download(y if in_b.value else x, x0.dtype, x0.shape)
        dsum = torch.zeros(eng.nb, dtype=torch.float64, device=eng.device)
        for _ in range(self.maxiter):
            rc = lib.jacobi_sweep(C.byref(eng.geom), C.byref(om), C.byref(el), x.data_ptr(), b.data_ptr(),
                                  y.data_ptr(), ws.data_ptr(), ws.numel(), eng.stream())
            lib.check(rc, "tvdk_jacobi_sweep")
            if self.tol is not None:
                # jacobi.py:118-131: error over the whole array, previous iterate kept on convergence
                rc = lib.norm_diff2(C.byref(eng.geom), y.data_ptr(), x.data_ptr(), dsum.data_ptr(), ws.data_ptr(),
                                    ws.numel(), eng.stream())
                lib.check(rc, "tvdk_norm_diff2")
                err = float(np.sqrt(dsum.cpu().numpy().sum())) / x0n
                if err < self.tol:
                    break
            x, y = y, x
        return eng.download(x, x0.dtype, x0.shape)


class CG(Solver):
    """Conjugate gradients with SciPy's recurrence (``cg.py:16-59`` + ``scipy.sparse.linalg.cg``)."""

    kind = _cabi.TVDK_SOLVER_CG

    def __call__(self, x0: np.ndarray, rhs: np.ndarray, h: Optional[float] = None) -> np.ndarray:
        if h is not None:
            raise NotImplementedError("only h = None (the value the TVD path uses) is implemented")
        if self.tol is None:
            raise TypeError("CG needs a numeric tol (SciPy rejects rtol=None); use tol=0.0 for a fixed count")
        import torch

        eng, om, el, _keep = self._engine(x0, rhs)
        lib, g = eng.lib, eng.geom
        nb = eng.nb
        x = eng.upload(x0.astype(eng.np_dtype, copy=False))
        b = eng.upload(rhs.astype(eng.np_dtype, copy=False))
        r = eng.empty_fields(1)
        p = [eng.empty_fields(1), eng.empty_fields(1)]
        ws = eng.op_workspace()
        s = eng.stream()
        sums = torch.zeros(nb, dtype=torch.float64, device=eng.device)
        coef = torch.zeros(nb, dtype=torch.float64, device=eng.device)

        def dev_sum():
            return sums.cpu().numpy().copy()

        # r = b - A x via the update kernel (p = x, alpha = 1, scratch "x" discarded); the whole
        # array is ONE linear system in the reference (flattened; batch axes are block-diagonal
        # in A), so per-problem sums are added up on the host
        r.copy_(b)
        scratch = eng.zero_fields(1)
        coef.fill_(1.0)
        lib.check(lib.cg_update(C.byref(g), C.byref(om), C.byref(el), x.data_ptr(), scratch.data_ptr(),
                                r.data_ptr(), coef.data_ptr(), sums.data_ptr(), ws.data_ptr(), ws.numel(), s),
                  "tvdk_cg_update")
        rho = float(dev_sum().sum())
        del scratch
        lib.check(lib.dot(C.byref(g), b.data_ptr(), b.data_ptr(), sums.data_ptr(), ws.data_ptr(), ws.numel(), s), "dot")
        bnrm = float(np.sqrt(dev_sum().sum()))
        if bnrm == 0.0:
            return np.zeros_like(x0)
        atol = max(0.0, float(self.tol) * bnrm)
        rho_prev = None
        for it in range(self.maxiter):
            if np.sqrt(rho) < atol:
                break
            beta = 0.0 if it == 0 else rho / rho_prev
            coef.fill_(beta)
            lib.check(lib.cg_pq(C.byref(g), C.byref(om), C.byref(el), p[(it & 1) ^ 1].data_ptr(), r.data_ptr(),
                                p[it & 1].data_ptr(), coef.data_ptr(), 1 if it == 0 else 0, sums.data_ptr(),
                                ws.data_ptr(), ws.numel(), s), "tvdk_cg_pq")
            alpha = rho / float(dev_sum().sum())
            coef.fill_(alpha)
            lib.check(lib.cg_update(C.byref(g), C.byref(om), C.byref(el), p[it & 1].data_ptr(), x.data_ptr(),
                                    r.data_ptr(), coef.data_ptr(), sums.data_ptr(), ws.data_ptr(), ws.numel(), s),
                      "tvdk_cg_update")
            rho_prev, rho = rho, float(dev_sum().sum())
        return eng.download(x, x0.dtype, x0.shape)


TVDK_SOLVER_MG = 2  # host-composed (darsia_b200/mg.py); not a tvdk_solve kind


class MG(Solver):
    """V-cycle multigrid with Jacobi smoothing (``mg.py:13-270``), scalar or per-voxel coefficients.

    Same constructor as the reference (``mg.py:25-35``).  The cycle runs on the GPU (``darsia_b200/mg.py``:
    sweeps, residual, restriction, prolongation are libtvdk kernels; the recursion is host logic).
    ``heterogeneous`` is decided here, from the constructor's coefficients, and never revisited (``mg.py:70-73``):
    per-voxel coefficients handed to an ``MG()`` built without arrays raise the reference's ``ValueError``.
    As in the reference, a heterogeneous solve leaves the block-averaged arrays in ``mass_coeff`` /
    ``diffusion_coeff`` (``mg.py:134-149, 170-192``).
    """

    kind = TVDK_SOLVER_MG

    def __init__(
        self,
        depth: int = 2,
        smoother_iterations: int = 5,
        maxiter: int = 100,
        tol: Optional[float] = None,
        dim: int = 2,
        mass_coeff: Optional[Coeff] = None,
        diffusion_coeff: Optional[Coeff] = None,
        verbose=False,
    ) -> None:
        super().__init__(maxiter=maxiter, tol=tol, dim=dim, mass_coeff=mass_coeff, diffusion_coeff=diffusion_coeff,
                         verbose=verbose)
        self.depth = depth
        self.smoother_iterations = smoother_iterations
        self.heterogeneous = isinstance(mass_coeff, np.ndarray) or isinstance(diffusion_coeff, np.ndarray)

    def __call__(self, x0: np.ndarray, rhs: np.ndarray) -> np.ndarray:
        from ..engine import Engine
        from ..mg import MGHierarchy

        assert x0.shape == rhs.shape
        if not np.issubdtype(x0.dtype, np.floating):
            raise TypeError("float arrays only")
        if x0.ndim > self.dim and self.tol is not None:
            # with a tolerance every trailing-axis slice has its own stopping cycle
            out = np.empty_like(x0)
            for idx in np.ndindex(x0.shape[self.dim:]):
                sel = (slice(None),) * self.dim + idx
                out[sel] = self(np.ascontiguousarray(x0[sel]), np.ascontiguousarray(rhs[sel]))
            return out
        dtype = np.float32 if x0.dtype == np.float32 else np.float64
        nb = int(np.prod(x0.shape[self.dim:], dtype=np.int64))
        eng = Engine(x0.shape[: self.dim], nb=nb, dtype=dtype)
        hier = MGHierarchy(eng, self.depth, self.smoother_iterations, self.mass_coeff, self.diffusion_coeff,
                           heterogeneous=self.heterogeneous, full_shape=x0.shape)
        x = eng.upload(x0.astype(dtype, copy=False))
        b = eng.upload(rhs.astype(dtype, copy=False))
        res = hier.solve(x, b, self.maxiter, self.tol)
        out = eng.download(res, x0.dtype, x0.shape)
        if hier.cycles:  # restrict_parameters / prolongate_parameters overwrote the solver's arrays (mg.py:134-149, 170-192)
            if "omega" in hier.maps:
                self.mass_coeff = hier.final_coefficient("omega", np.shape(self.mass_coeff)).astype(np.asarray(self.mass_coeff).dtype)
            if "ell" in hier.maps:
                self.diffusion_coeff = hier.final_coefficient("ell", np.shape(self.diffusion_coeff)).astype(
                    np.asarray(self.diffusion_coeff).dtype)
        return out
