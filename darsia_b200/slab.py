"""Slab decomposition of ONE large 3-D (or 2-D) volume across ranks (BASELINE configs 4/5).

The volume is cut along axis 0 into contiguous slabs, one per rank (``torch.distributed``: NCCL
on GPUs).  Every stencil application is preceded by a halo exchange of whole planes (planes of
axis 0 are contiguous in the device layout, so a halo is a plain send/recv of a slice -- no
pack kernel), and every dot product / norm is an all-reduce of one or two doubles.  All global
boundary rules are evaluated on global plane indices inside the kernels (``tvdk_geom_t``:
``g0_off``, ``g0_n``), so the decomposed run performs the same arithmetic per voxel as the
single-GPU run; only the summation order of the reductions differs (about 1e-16 relative).

Reference semantics: ``split_bregman_tvd`` (restoration/split_bregman_tvd.py:14-177) with
``Jacobi`` (jacobi.py:74-133, incl. the ``tol`` branch) or ``CG`` (cg.py:16-59 + SciPy's recurrence).
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
    def exchange(self, views: List[Tuple[torch.Tensor, int, int, int]], wait: bool = True):
        """Halo exchange.  ``views``: (field, component, n_up, n_down) -- the last ``n_up`` owned
        planes go to the upper neighbour's ghost planes below its range, the first ``n_down``
        owned planes go to the lower neighbour's ghost planes above its range.
        With ``wait=False`` the transfers are only enqueued (they start once everything enqueued on the
        current stream so far has finished) and the request list is returned: kernels launched
        afterwards overlap with them; ``finish`` makes the current stream wait for the data."""
        if self.world == 1:
            return []
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
        reqs = dist.batch_isend_irecv(ops) if ops else []
        if wait:
            self.finish(reqs)
            return []
        return reqs

    @staticmethod
    def finish(reqs) -> None:
        for req in reqs:
            req.wait()

    def sub_geom(self, z0: int, z1: int) -> "_cabi.Geom":
        """The slab's geometry restricted to the owned planes [z0, z1) (local indices)."""
        g = _cabi.Geom.from_buffer_copy(self.geom)
        g.z_begin, g.z_end = int(z0), int(z1)
        return g

    def allreduce(self, t: torch.Tensor) -> torch.Tensor:
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t


class SlabSolver:
    """Device-resident state of one slab-decomposed ``split_bregman_tvd`` run.

    ``setup`` uploads this rank's planes and exchanges the static halos; ``iterate(n)`` enqueues
    ``n`` outer iterations (halo exchanges + kernels + all-reduces) without any host
    synchronisation unless ``eps`` or a positive CG tolerance asks for one; ``result`` downloads
    the owned planes.  ``bench.py`` times ``iterate``.
    """

    HB = 4  # planes per boundary launch of the overlapped fused route (>= the 2 planes sent down)

    def __init__(self, global_shape, dim, dtype, mu, omega, ell, isotropic, solver, eps=None, group=None,
                 device=None):
        rank = dist.get_rank(group) if dist.is_initialized() else 0
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        if len(global_shape) != dim:
            raise AssertionError("slab decomposition handles one volume (no batch axes)")
        self.spec = _solver_spec(solver)
        if ell is None:
            ell = 2 * mu
        _check_positive("ell", ell)
        self.dim, self.iso, self.eps = dim, 1 if isotropic else 0, eps
        self.cdtype = np.dtype(dtype)
        self.eng = eng = SlabEngine(global_shape, rank, world, dtype=self.cdtype, device=device, group=group)
        self._coefs = [self._coef(v, n) for v, n in ((mu, "mu"), (omega, "omega"), (ell, "ell"))]
        self.fused = self.spec.kind == _cabi.TVDK_SOLVER_JACOBI and self.spec.maxiter == 1 and self.spec.tol is None
        self.cur = 0
        self.incs: List[torch.Tensor] = []
        self.done = False
        self.iters = 0
        self._pending = None  # halo transfers of the current iterate still in flight (fused route)

    def _coef(self, value, name):
        eng = self.eng
        if np.ndim(value) == 0:
            return eng.coef(float(value), name=name)
        fld = eng.upload_owned(np.asarray(value, dtype=self.cdtype))
        eng.exchange([(fld, 0, GL, GH)])
        return eng.coef(fld, name=name)

    def setup(self, img_owned: np.ndarray) -> None:
        eng, D = self.eng, self.dim
        lib, g = eng.lib, eng.geom
        self.f = eng.upload_owned(np.asarray(img_owned, dtype=self.cdtype))
        eng.exchange([(self.f, 0, GL, GH)])
        self.u = self.f.clone()
        self.t = eng.zero_fields(D)
        ws = eng.op_workspace()
        self.sums = torch.zeros(8, dtype=torch.float64, device=eng.device)  # 3 launches x up to 2 sums
        self._tot = torch.zeros(2, dtype=torch.float64, device=eng.device)
        # Boundary-first splitting of the CG kernels: measured +12 % on 2 GPUs, -4 % on 8 (1024^3: the NCCL
        # transfers then share the SMs with the interior launch and every kernel pays two more small launches),
        # +2..3 % for the fused Jacobi-1 kernel on 2 and on 8 GPUs (profiles/r1_slab_*).
        self._split_cg = eng.world <= 4
        with device_guard(eng.device):  # ||img|| over all ranks (split_bregman_tvd.py:58)
            lib.check(lib.dot(C.byref(g), self.f.data_ptr(), self.f.data_ptr(), self.sums.data_ptr(), ws.data_ptr(),
                              ws.numel(), eng.stream()), "tvdk_dot")
        self.nrm2 = eng.allreduce(self.sums[:1].clone())
        if self.fused:
            self.bufs = [(self.u, self.t), (eng.empty_fields(1), eng.zero_fields(D))]
        elif self.spec.kind == _cabi.TVDK_SOLVER_JACOBI:
            self.rhs, self.xa, self.xb = eng.empty_fields(1), eng.zero_fields(1), eng.zero_fields(1)
        else:
            self.x, self.r = eng.zero_fields(1), eng.zero_fields(1)
            self.p = [eng.zero_fields(1), eng.zero_fields(1)]
            self.rho, self.rho_prev, self.pq, self.alpha, self.beta = (
                torch.zeros(1, dtype=torch.float64, device=eng.device) for _ in range(5))

    def _split(self, launch, halo, nsum: int = 1, lo_planes: int = 0, hi_planes: int = 0, overlap: bool = True):
        """Boundary-first overlap of one kernel with the halo exchange of what it produces.

        ``launch(geom_ptr, sum_ptr)`` enqueues the kernel on the owned planes of ``geom``; it is called for
        the few planes at either end of the slab that the neighbours need, then the transfers ``halo``
        are enqueued (NCCL starts them once those small launches are done) and the interior launch
        runs meanwhile.  The partial sums of the launches are added into ``self.sums[:nsum]``.
        Returns the pending requests (``SlabEngine.finish`` before the ghosts are read)."""
        eng, sums = self.eng, self.sums
        nlo = max(self.HB, lo_planes) if eng.lo is not None else 0
        nhi = max(self.HB, hi_planes) if eng.hi is not None else 0

        def run(z0, z1, slot):
            gg = eng.sub_geom(z0, z1)
            launch(C.byref(gg), sums[slot * nsum:(slot + 1) * nsum].data_ptr())

        if overlap and eng.world > 1 and eng.nown >= nlo + nhi + 1:
            if not (nlo and nhi):  # the slot of the missing end must not carry an old value
                sums[nsum:3 * nsum].zero_()
            if nlo:
                run(eng.zb, eng.zb + nlo, 1)
            if nhi:
                run(eng.ze - nhi, eng.ze, 2)
            reqs = eng.exchange(halo, wait=False)
            run(eng.zb + nlo, eng.ze - nhi, 0)
            torch.sum(sums[: 3 * nsum].view(3, nsum), dim=0, out=self._tot[:nsum])
            sums[:nsum].copy_(self._tot[:nsum])
        else:
            run(eng.zb, eng.ze, 0)
            reqs = eng.exchange(halo, wait=False)
        return reqs

    def iterate(self, n: int) -> int:
        """Enqueue up to ``n`` outer iterations; returns how many were run (eps may stop early)."""
        eng, spec, D, iso = self.eng, self.spec, self.dim, self.iso
        lib, g = eng.lib, eng.geom
        (mu_c, _), (om_c, _), (el_c, _) = self._coefs
        bm, bo, be, bg = C.byref(mu_c), C.byref(om_c), C.byref(el_c), C.byref(g)
        ws = eng.op_workspace()
        wsp, wsn = ws.data_ptr(), ws.numel()
        sums, tcs = self.sums, eng.batch_elems
        call = lib.check
        ran = 0
        while ran < n and not self.done:
            with device_guard(eng.device):
                s_ = eng.stream()
                if self.fused:
                    # Boundary-first overlap (``_split``): the planes the neighbours need (2 at the lower, 1 at
                    # the upper end) come from two small launches, the interior launch hides their transfer.
                    (ui, ti), (uo, to) = self.bufs[self.cur], self.bufs[1 - self.cur]
                    if self._pending is None:  # first iteration (or after reset): ghosts of the input
                        eng.exchange([(ui, 0, 1, 2)] + [(ti, d, 1, 1) for d in range(D)])
                    else:
                        eng.finish(self._pending)
                    self._pending = self._split(
                        lambda gp, sp: call(lib.fused_jacobi1(gp, iso, bm, bo, be, self.f.data_ptr(), ui.data_ptr(),
                                                              uo.data_ptr(), ti.data_ptr(), to.data_ptr(), tcs, sp, wsp,
                                                              wsn, s_), "tvdk_fused_jacobi1"),
                        [(uo, 0, 1, 2)] + [(to, d, 1, 1) for d in range(D)])
                    self.cur = 1 - self.cur
                elif spec.kind == _cabi.TVDK_SOLVER_JACOBI:
                    u, t = self.u, self.t
                    eng.exchange([(t, d, 1, 1) for d in range(D)])
                    call(lib.rhs(bg, iso, bm, bo, be, self.f.data_ptr(), t.data_ptr(), tcs, self.rhs.data_ptr(), wsp, wsn,
                                 s_), "tvdk_rhs")
                    xin = u
                    x0n2 = None
                    for k in range(spec.maxiter):
                        xout = self.xb if xin is self.xa else self.xa
                        eng.exchange([(xin, 0, 1, 1)])
                        call(lib.jacobi_sweep(bg, bo, be, xin.data_ptr(), self.rhs.data_ptr(), xout.data_ptr(), wsp, wsn,
                                              s_), "tvdk_jacobi_sweep")
                        if spec.tol is not None:
                            # jacobi.py:118-131 over the whole volume: err = ||x_new - x|| / ||x0||; on convergence the
                            # PREVIOUS iterate is kept (one host sync per sweep, only with a tolerance)
                            if x0n2 is None:
                                call(lib.dot(bg, xin.data_ptr(), xin.data_ptr(), sums.data_ptr(), wsp, wsn, s_), "tvdk_dot")
                                x0n2 = float(eng.allreduce(sums[0:1].clone()).item())
                            call(lib.norm_diff2(bg, xout.data_ptr(), xin.data_ptr(), sums.data_ptr(), wsp, wsn, s_),
                                 "tvdk_norm_diff2")
                            d2 = float(eng.allreduce(sums[0:1].clone()).item())
                            with np.errstate(divide="ignore", invalid="ignore"):
                                if np.sqrt(d2) / np.sqrt(x0n2) < spec.tol:
                                    break
                        xin = xout
                    if xin is u:  # converged before the first accepted sweep: u+ = u (the shrink kernel needs a copy)
                        self.xa.copy_(u)
                        xin = self.xa
                    eng.exchange([(xin, 0, 0, 1)])
                    call(lib.shrink(bg, iso, bm, be, xin.data_ptr(), u.data_ptr(), u.data_ptr(), t.data_ptr(),
                                    t.data_ptr(), tcs, sums.data_ptr(), wsp, wsn, s_), "tvdk_shrink")
                else:
                    u, t, x, r, p = self.u, self.t, self.x, self.r, self.p
                    rho, rho_prev, pq, alpha, beta = self.rho, self.rho_prev, self.pq, self.alpha, self.beta
                    # every kernel is split boundary-first (``_split``) so that the halo exchange of what it
                    # produces overlaps its interior launch; the all-reduce that follows each kernel also
                    # orders the transfers before the next kernel's ghost reads
                    if self._pending is None:
                        eng.exchange([(u, 0, 1, 1)] + [(t, d, 1, 1) for d in range(D)])
                    else:
                        eng.finish(self._pending)
                    reqs = self._split(
                        lambda gp, sp: call(lib.cg_init(gp, iso, bm, bo, be, self.f.data_ptr(), t.data_ptr(), tcs,
                                                        u.data_ptr(), x.data_ptr(), r.data_ptr(), sp, wsp, wsn, s_),
                                            "tvdk_cg_init"),
                        [(r, 0, 1, 1)], nsum=2, overlap=self._split_cg)
                    eng.allreduce(sums[0:2])
                    bnrm = torch.sqrt(sums[0:1])
                    atol = float(spec.tol) * bnrm
                    rho.copy_(sums[1:2])
                    stop = False
                    if spec.tol > 0.0 or self.iters == 0:  # host check only when a tolerance can fire
                        zero_rhs = bool((bnrm == 0).item())
                        if zero_rhs:  # scipy cg returns b (all zeros) for a zero right-hand side
                            x.zero_()
                        stop = zero_rhs or bool((torch.sqrt(rho) < atol).item())
                    updated = False
                    for i in range(spec.maxiter):
                        if stop:
                            break
                        eng.finish(reqs)  # ghosts of r
                        if i > 0:
                            torch.div(rho, rho_prev, out=beta)
                        pin, pout = p[(i & 1) ^ 1], p[i & 1]
                        reqs = self._split(
                            lambda gp, sp: call(lib.cg_pq(gp, bo, be, pin.data_ptr(), r.data_ptr(), pout.data_ptr(),
                                                          beta.data_ptr(), 1 if i == 0 else 0, sp, wsp, wsn, s_),
                                                "tvdk_cg_pq"),
                            [(pout, 0, 1, 1)], overlap=self._split_cg)
                        pq.copy_(eng.allreduce(sums[0:1]))
                        torch.div(rho, pq, out=alpha)
                        eng.finish(reqs)  # ghosts of p
                        last = i == spec.maxiter - 1
                        reqs = self._split(
                            lambda gp, sp: call(lib.cg_update(gp, bo, be, pout.data_ptr(), x.data_ptr(), r.data_ptr(),
                                                              alpha.data_ptr(), sp, wsp, wsn, s_), "tvdk_cg_update"),
                            [(r, 0, 1, 1)] + ([(x, 0, 0, 1)] if (last or spec.tol > 0.0) else []), overlap=self._split_cg)
                        updated = True
                        rho_prev.copy_(rho)
                        rho.copy_(eng.allreduce(sums[0:1]))
                        if spec.tol > 0.0:
                            stop = bool((torch.sqrt(rho) < atol).item())
                    eng.finish(reqs)
                    if not updated:
                        eng.exchange([(x, 0, 0, 1)])  # stopped before any update: x = u still needs its upper ghost
                    self._pending = self._split(
                        lambda gp, sp: call(lib.shrink(gp, iso, bm, be, x.data_ptr(), u.data_ptr(), u.data_ptr(),
                                                       t.data_ptr(), t.data_ptr(), tcs, sp, wsp, wsn, s_), "tvdk_shrink"),
                        [(u, 0, 1, 1)] + [(t, d, 1, 1) for d in range(D)], lo_planes=1, overlap=self._split_cg)
            inc2 = eng.allreduce(sums[0:1].clone())
            self.incs.append(torch.sqrt(inc2 / self.nrm2))
            ran += 1
            self.iters += 1
            if self.eps is not None:  # split_bregman_tvd.py:173-175 (one host sync per iteration, only with eps)
                if float(self.incs[-1].item()) < self.eps:
                    self.done = True
        return ran

    def reset(self) -> None:
        """Back to the initial state (u = img, t = 0) without re-uploading."""
        if self._pending is not None:
            self.eng.finish(self._pending)
            self._pending = None
        if self.fused:
            self.cur = 0
        self.u.copy_(self.f)
        self.t.zero_()
        self.incs, self.done, self.iters = [], False, 0

    def result_field(self) -> torch.Tensor:
        if self._pending is not None:  # the last iterate's halo transfers: nothing reads them, but let them land
            self.eng.finish(self._pending)
            self._pending = None
        return self.bufs[self.cur][0] if self.fused else self.u

    def result(self, out_dtype=None) -> np.ndarray:
        return self.eng.download_owned(self.result_field(), out_dtype)

    def history(self) -> List[float]:
        return [float(v.item()) for v in self.incs]


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
        mu, omega, ell: scalars, or per-voxel maps given as this rank's owned planes.
        Other arguments as ``split_bregman_tvd``; ``solver`` is ``Jacobi(maxiter=K[, tol=...])`` or
        ``CG(maxiter=m, tol=...)``.
    Returns this rank's owned planes of the result (and, with ``return_info``, the relative increments).
    """
    if not isinstance(img_owned, np.ndarray) or img_owned.dtype not in (np.float32, np.float64):
        raise TypeError("float32/float64 ndarray expected")
    cdtype = img_owned.dtype if precision is None else np.dtype({"fp64": np.float64, "fp32": np.float32}[precision])
    sol = SlabSolver(global_shape, dim, cdtype, mu, omega, ell, isotropic, solver, eps=eps, group=group, device=device)
    sol.setup(img_owned)
    sol.iterate(max_num_iter)
    out = sol.result(img_owned.dtype)
    return (out, sol.history()) if return_info else out
