"""Slab decomposition of ONE large 3-D (or 2-D) volume across ranks (BASELINE configs 4/5).

The volume is cut along axis 0 into contiguous slabs, one per rank (``torch.distributed``: NCCL
on GPUs).  Every stencil application is preceded by a halo exchange of whole planes (planes of
axis 0 are contiguous in the device layout, so a halo is a plain send/recv of a slice -- no
pack kernel), and every dot product / norm is an all-reduce of one or two doubles.  All global
boundary rules are evaluated on global plane indices inside the kernels (``tvdk_geom_t``:
``g0_off``, ``g0_n``), so the decomposed run performs the same arithmetic per voxel as the
single-GPU run; only the summation order of the reductions differs (about 1e-16 relative).

Reference semantics: ``split_bregman_tvd`` (restoration/split_bregman_tvd.py:14-177) with
``Jacobi`` (jacobi.py:74-133, without ``tol``) or ``CG`` (cg.py:16-59 + SciPy's recurrence).
The reference itself has no distributed code; this is row (e) of SURVEY.md section 8.
"""

from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist

from . import _cabi
from ._runtime import device_guard
from .engine import Engine
from .restoration.split_bregman_tvd import _check_positive, _solver_spec

GL, GH = 1, 2  # ghost planes below / above the owned range


def slab_range(n0: int, rank: int, world: int) -> Tuple[int, int]:
    """Owned global planes [s, e) of ``rank``: as even as possible, lower ranks get the rest."""
    base, rem = divmod(n0, world)
    s = rank * base + min(rank, rem)
    return s, s + base + (1 if rank < rem else 0)


class SlabEngine(Engine):
    """An ``Engine`` whose fields are one slab (owned planes + ghosts) of a global volume."""

    def __init__(self, global_shape: Sequence[int], rank: int, world: int, dtype=np.float64, device=None,
                 group=None, tile=(0, 0, 0)) -> None:
        self.global_shape = tuple(int(n) for n in global_shape)
        self.rank, self.world, self.group = rank, world, group
        n0 = self.global_shape[0]
        if n0 < world:
            raise ValueError(f"cannot cut {n0} planes into {world} slabs")
        self.s, self.e = slab_range(n0, rank, world)
        self.nown = self.e - self.s
        if world > 1 and min(slab_range(n0, r, world)[1] - slab_range(n0, r, world)[0] for r in range(world)) < GH:
            raise ValueError("every slab needs at least 2 planes")
        local = (GL + self.nown + GH,) + self.global_shape[1:]
        super().__init__(local, nb=1, dtype=dtype, device=device, tile=tile,
                         slab=(self.s - GL, n0, GL, GL + self.nown))
        self.zb, self.ze = GL, GL + self.nown
        self.lo = rank - 1 if rank > 0 else None  # neighbour owning the planes below
        self.hi = rank + 1 if rank < world - 1 else None

    # -- layout -----------------------------------------------------------------------------
    def upload_owned(self, owned: np.ndarray) -> torch.Tensor:
        """Host array of the owned planes -> slab field (ghost planes zero)."""
        assert owned.shape == (self.nown,) + self.global_shape[1:], owned.shape
        full = np.zeros((GL + self.nown + GH,) + self.global_shape[1:], dtype=owned.dtype)
        full[self.zb:self.ze] = owned
        return self.upload(full)

    def download_owned(self, field: torch.Tensor, out_dtype=None) -> np.ndarray:
        full = self.download(field, out_dtype, shape=self.spatial)
        return np.ascontiguousarray(full[self.zb:self.ze])

    def planes(self, field: torch.Tensor, z0: int, z1: int, comp: int = 0) -> torch.Tensor:
        """Contiguous view of local planes [z0, z1) of component ``comp`` of a field tensor."""
        base = comp * self.batch_elems
        return field[base + z0 * self.plane_elems: base + z1 * self.plane_elems]

    # -- communication ----------------------------------------------------------------------
    def exchange(self, views: List[Tuple[torch.Tensor, int, int, int]]) -> None:
        """Halo exchange.  ``views``: (field, component, n_up, n_down) -- the last ``n_up`` owned
        planes go to the upper neighbour's ghost planes below its range, the first ``n_down``
        owned planes go to the lower neighbour's ghost planes above its range."""
        if self.world == 1:
            return
        ops = []
        for field, comp, n_up, n_down in views:
            assert n_up <= GL and n_down <= GH
            if n_up:
                if self.hi is not None:
                    ops.append(dist.P2POp(dist.isend, self.planes(field, self.ze - n_up, self.ze, comp), self.hi, self.group))
                if self.lo is not None:
                    ops.append(dist.P2POp(dist.irecv, self.planes(field, self.zb - n_up, self.zb, comp), self.lo, self.group))
            if n_down:
                if self.lo is not None:
                    ops.append(dist.P2POp(dist.isend, self.planes(field, self.zb, self.zb + n_down, comp), self.lo, self.group))
                if self.hi is not None:
                    ops.append(dist.P2POp(dist.irecv, self.planes(field, self.ze, self.ze + n_down, comp), self.hi, self.group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()

    def allreduce(self, t: torch.Tensor) -> torch.Tensor:
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t


def split_bregman_tvd_slab(
    img_owned: np.ndarray,
    global_shape: Sequence[int],
    mu=1.0,
    omega=1.0,
    ell=None,
    dim: int = 3,
    max_num_iter: int = 100,
    eps: Optional[float] = None,
    isotropic: bool = False,
    solver=None,
    group=None,
    device=None,
    precision: Optional[str] = None,
    return_info: bool = False,
):
    """``split_bregman_tvd`` for a volume distributed over the ranks of ``group``.

    Args:
        img_owned: this rank's planes ``[s, e) = slab_range(global_shape[0], rank, world)`` of the image.
        global_shape: shape of the whole volume.
        mu, omega, ell: scalars (per-voxel maps: pass this rank's owned planes).
        Other arguments as ``split_bregman_tvd``; ``solver`` is ``Jacobi(maxiter=K)`` (no tol) or
        ``CG(maxiter=m, tol=...)``.
    Returns this rank's owned planes of the result (and the list of relative increments).
    """
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if len(global_shape) != dim:
        raise AssertionError("slab decomposition handles one volume (no batch axes)")
    if not isinstance(img_owned, np.ndarray) or img_owned.dtype not in (np.float32, np.float64):
        raise TypeError("float32/float64 ndarray expected")
    spec = _solver_spec(solver)
    if spec.kind == _cabi.TVDK_SOLVER_JACOBI and spec.tol is not None:
        raise NotImplementedError("Jacobi with tol is not implemented for slabs")
    if ell is None:
        ell = 2 * mu
    _check_positive("ell", ell)
    cdtype = img_owned.dtype if precision is None else np.dtype({"fp64": np.float64, "fp32": np.float32}[precision])
    eng = SlabEngine(global_shape, rank, world, dtype=cdtype, device=device, group=group)
    lib, g = eng.lib, eng.geom
    iso = 1 if isotropic else 0
    D = dim

    def coef(value, name):
        if np.ndim(value) == 0:
            return eng.coef(float(value), name=name)
        fld = eng.upload_owned(np.asarray(value, dtype=cdtype))
        eng.exchange([(fld, 0, GL, GH)])
        return eng.coef(fld, name=name)

    (mu_c, k1), (om_c, k2), (el_c, k3) = coef(mu, "mu"), coef(omega, "omega"), coef(ell, "ell")
    bm, bo, be, bg = C.byref(mu_c), C.byref(om_c), C.byref(el_c), C.byref(g)
    f = eng.upload_owned(img_owned)
    eng.exchange([(f, 0, GL, GH)])
    u = f.clone()
    t = eng.zero_fields(D)
    ws = eng.op_workspace()
    wsp, wsn = ws.data_ptr(), ws.numel()
    sums = torch.zeros(2, dtype=torch.float64, device=eng.device)
    tcs = eng.batch_elems

    def call(rc, what):
        lib.check(rc, what)

    # ||img|| over all ranks (split_bregman_tvd.py:58)
    with device_guard(eng.device):
        call(lib.dot(bg, f.data_ptr(), f.data_ptr(), sums.data_ptr(), wsp, wsn, eng.stream()), "tvdk_dot")
    nrm2 = eng.allreduce(sums[:1].clone())
    incs: List[torch.Tensor] = []
    hist: List[float] = []
    tviews = [(t, d, 1, 1) for d in range(D)]

    if spec.kind == _cabi.TVDK_SOLVER_JACOBI and spec.maxiter == 1:
        u2, t2 = eng.empty_fields(1), eng.zero_fields(D)
        bufs = [(u, t), (u2, t2)]
    elif spec.kind == _cabi.TVDK_SOLVER_JACOBI:
        rhs, xa, xb = eng.empty_fields(1), eng.zero_fields(1), eng.zero_fields(1)
    else:
        x, r = eng.zero_fields(1), eng.zero_fields(1)
        p = [eng.zero_fields(1), eng.zero_fields(1)]
        rho, rho_prev, pq, alpha, beta = (torch.zeros(1, dtype=torch.float64, device=eng.device) for _ in range(5))

    cur = 0
    done = False
    it = 0
    while it < max_num_iter and not done:
        with device_guard(eng.device):
            s_ = eng.stream()
            if spec.kind == _cabi.TVDK_SOLVER_JACOBI and spec.maxiter == 1:
                (ui, ti), (uo, to) = bufs[cur], bufs[1 - cur]
                eng.exchange([(ui, 0, 1, 2)] + [(ti, d, 1, 1) for d in range(D)])
                call(lib.fused_jacobi1(bg, iso, bm, bo, be, f.data_ptr(), ui.data_ptr(), uo.data_ptr(), ti.data_ptr(),
                                       to.data_ptr(), tcs, sums.data_ptr(), wsp, wsn, s_), "tvdk_fused_jacobi1")
                cur = 1 - cur
            elif spec.kind == _cabi.TVDK_SOLVER_JACOBI:
                eng.exchange(tviews)
                call(lib.rhs(bg, iso, bm, bo, be, f.data_ptr(), t.data_ptr(), tcs, rhs.data_ptr(), wsp, wsn, s_), "tvdk_rhs")
                xin = u
                for k in range(spec.maxiter):
                    xout = xa if k % 2 == 0 else xb
                    eng.exchange([(xin, 0, 1, 1)])
                    call(lib.jacobi_sweep(bg, bo, be, xin.data_ptr(), rhs.data_ptr(), xout.data_ptr(), wsp, wsn, s_),
                         "tvdk_jacobi_sweep")
                    xin = xout
                eng.exchange([(xin, 0, 0, 1)])
                call(lib.shrink(bg, iso, bm, be, xin.data_ptr(), u.data_ptr(), u.data_ptr(), t.data_ptr(), t.data_ptr(),
                                tcs, sums.data_ptr(), wsp, wsn, s_), "tvdk_shrink")
            else:
                eng.exchange([(u, 0, 1, 1)] + tviews)
                call(lib.cg_init(bg, iso, bm, bo, be, f.data_ptr(), t.data_ptr(), tcs, u.data_ptr(), x.data_ptr(),
                                 r.data_ptr(), sums.data_ptr(), wsp, wsn, s_), "tvdk_cg_init")
                eng.allreduce(sums)
                bnrm = torch.sqrt(sums[0:1])
                atol = float(spec.tol) * bnrm
                rho.copy_(sums[1:2])
                stop = False
                if spec.tol > 0.0 or it == 0:  # host check only when a tolerance can fire (or once: zero rhs)
                    stop = bool((bnrm == 0).item()) or bool((torch.sqrt(rho) < atol).item())
                for i in range(spec.maxiter):
                    if stop:
                        break
                    eng.exchange([(r, 0, 1, 1)])
                    if i > 0:
                        torch.div(rho, rho_prev, out=beta)
                    call(lib.cg_pq(bg, bo, be, p[(i & 1) ^ 1].data_ptr(), r.data_ptr(), p[i & 1].data_ptr(),
                                   beta.data_ptr(), 1 if i == 0 else 0, sums.data_ptr(), wsp, wsn, s_), "tvdk_cg_pq")
                    pq.copy_(eng.allreduce(sums[0:1]))
                    torch.div(rho, pq, out=alpha)
                    eng.exchange([(p[i & 1], 0, 1, 1)])
                    call(lib.cg_update(bg, bo, be, p[i & 1].data_ptr(), x.data_ptr(), r.data_ptr(), alpha.data_ptr(),
                                       sums.data_ptr(), wsp, wsn, s_), "tvdk_cg_update")
                    rho_prev.copy_(rho)
                    rho.copy_(eng.allreduce(sums[0:1]))
                    if spec.tol > 0.0:
                        stop = bool((torch.sqrt(rho) < atol).item())
                eng.exchange([(x, 0, 0, 1)])
                call(lib.shrink(bg, iso, bm, be, x.data_ptr(), u.data_ptr(), u.data_ptr(), t.data_ptr(), t.data_ptr(),
                                tcs, sums.data_ptr(), wsp, wsn, s_), "tvdk_shrink")
        inc2 = eng.allreduce(sums[0:1].clone())
        incs.append(torch.sqrt(inc2 / nrm2))
        it += 1
        if eps is not None:  # split_bregman_tvd.py:173-175 (one host sync per iteration, only with eps)
            v = float(incs[-1].item())
            if v < eps:
                done = True
    res_field = bufs[cur][0] if (spec.kind == _cabi.TVDK_SOLVER_JACOBI and spec.maxiter == 1) else u
    out = eng.download_owned(res_field, img_owned.dtype)
    if return_info:
        hist = [float(v.item()) for v in incs]
        return out, hist
    return out
