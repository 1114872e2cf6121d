"""Slab decomposition of ONE large volume over several GPUs, driven by ``tvdk_solve_slab``
(``include/tvdk.h``): the whole outer split-Bregman loop, the halo exchange and every dot-product
all-reduce run on the devices -- kernels push boundary planes and partial sums straight into the peers'
memory over NVLink (CUDA peer access / CUDA IPC) and a one-warp kernel adds the sums in rank order.
No NCCL call and no Python on the data path; ``torch.distributed`` is only used once, to trade the
64-byte IPC handles of the arenas.

Reference semantics: ``split_bregman_tvd`` (restoration/split_bregman_tvd.py:14-177) with ``Jacobi``
(jacobi.py:74-133, incl. the ``tol`` branch) or ``CG`` (cg.py:16-59 + SciPy's recurrence).  The reference
itself has no distributed code; this is row (e) of SURVEY.md section 8 (BASELINE configs 4 and 5).

Two ways to run a decomposition of ``world`` slabs:

* ``SlabGroup.distributed(...)``: one process per GPU (``torchrun``), this process drives ONE slab;
* ``SlabGroup.local(world, devices=...)``: this process drives ALL slabs in lock step -- on several GPUs, or
  on one device (tests: the slab geometry, the pushes and the ordered sums run on a single GPU; the ranks are
  ordered by stream order instead of flags, nothing ever waits on a concurrently running kernel).
"""

from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _cabi
from ._runtime import device_guard, resolve_device, stream_ptr
from .restoration.split_bregman_tvd import _check_positive, _solver_spec
from .slab import slab_range

GL, GH = _cabi.TVDK_SLAB_GL, _cabi.TVDK_SLAB_GH
_TORCH = {np.dtype(np.float64): torch.float64, np.dtype(np.float32): torch.float32}
_CODE = {np.dtype(np.float64): _cabi.TVDK_F64, np.dtype(np.float32): _cabi.TVDK_F32}


def _align(x: int, a: int = 256) -> int:
    return (x + a - 1) // a * a


class _Rank:
    """Arena, geometry and problem descriptor of one slab."""

    def __init__(self, lib, global_shape, rank, world, nown, cdtype, device, spec, iso, eps, coef_kinds):
        self.lib, self.rank, self.world, self.device = lib, rank, world, device
        self.shape = tuple(global_shape)
        self.dim = len(self.shape)
        self.cdtype = np.dtype(cdtype)
        self.nown = list(nown)
        self.s = sum(self.nown[:rank])
        cap = max(self.nown)
        g = _cabi.Geom()
        g.dim, g.dtype = self.dim, _CODE[self.cdtype]
        local = (GL + cap + GH,) + self.shape[1:]
        for a in range(3):
            g.n[a] = local[a] if a < self.dim else 1
        g.pitch = lib.pitch(g.dtype, self.shape[-1])
        g.nb = 1
        g.g0_off, g.g0_n, g.z_begin, g.z_end = self.s - GL, self.shape[0], GL, GL + self.nown[rank]
        self.geom = g
        self.field_elems = int(lib.field_elems(C.byref(g)))
        if self.field_elems <= 0:
            raise ValueError(lib.last_error().decode())
        self.plane_elems = self.field_elems // local[0]
        W = self.cdtype.itemsize
        fb = _align(self.field_elems * W)
        # arena layout (identical offsets on every rank): header | f | u | t (dim) | maps | stats | history | workspace
        p = _cabi.Problem()
        p.geom = g
        p.isotropic = 1 if iso else 0
        p.solver_kind, p.solver_maxiter = spec.kind, int(spec.maxiter)
        p.solver_tol = -1.0 if spec.tol is None else float(spec.tol)
        p.eps = -1.0 if eps is None else float(eps)
        p.fused, p.check_every = 1, 8
        self.ws_bytes = int(lib.slab_workspace_bytes(C.byref(p)))
        if self.ws_bytes == 0:
            raise ValueError(lib.last_error().decode())
        off = _align(int(lib.slab_header_bytes()))
        self.off = {}
        for name, n in (("f", 1), ("u", 1), ("t", self.dim)):
            self.off[name] = off
            off += n * fb
        for name in coef_kinds:  # per-voxel maps
            self.off[name] = off
            off += fb
        self.off["stats"] = off
        off += _align(C.sizeof(_cabi.Stat))
        self.hist_cap = 4096
        self.off["hist"] = off
        off += _align(8 * self.hist_cap)
        self.off["ws"] = off
        off += _align(self.ws_bytes)
        self.arena_bytes = off
        self.field_bytes = fb
        self.arena = torch.zeros(self.arena_bytes, dtype=torch.uint8, device=device)
        self.base = self.arena.data_ptr()
        p.f, p.u, p.t = self.base + self.off["f"], self.base + self.off["u"], self.base + self.off["t"]
        p.t_comp_stride = fb // W
        p.workspace, p.workspace_bytes = self.base + self.off["ws"], self.ws_bytes
        p.stats = self.base + self.off["stats"]
        p.history = self.base + self.off["hist"]
        self.problem = p
        self.comm = _cabi.SlabComm()
        self.comm.world, self.comm.rank = world, rank
        for r in range(world):
            self.comm.nown[r] = self.nown[r]
        self.comm.arena_bytes = self.arena_bytes
        self.comm.timeout_ms = 20000

    # views ------------------------------------------------------------------------------------
    def field(self, name: str, comp: int = 0) -> torch.Tensor:
        W = self.cdtype.itemsize
        o = self.off[name] + comp * self.field_bytes
        return self.arena[o:o + self.field_elems * W].view(_TORCH[self.cdtype])

    def owned_geom(self) -> _cabi.Geom:
        """Geometry of the owned planes alone (for the layout kernels)."""
        g = _cabi.Geom.from_buffer_copy(self.geom)
        n = self.nown[self.rank]
        g.n[0] = n
        g.g0_off, g.g0_n, g.z_begin, g.z_end = 0, n, 0, n
        return g

    def stream(self) -> int:
        return stream_ptr(self.lib, self.device)

    def put_owned(self, name: str, owned, comp: int = 0) -> None:
        """Host (or device) array of this rank's owned planes -> the field's owned planes (H2D + layout kernel;
        no padded host copy)."""
        n = self.nown[self.rank]
        if tuple(owned.shape) != (n,) + self.shape[1:]:
            raise AssertionError(f"rank {self.rank}: expected owned planes of shape {(n,) + self.shape[1:]}, got {tuple(owned.shape)}")
        if isinstance(owned, torch.Tensor):
            src = owned.to(self.device, non_blocking=True).contiguous().reshape(-1)
        else:
            src = torch.from_numpy(np.ascontiguousarray(owned).reshape(-1)).to(self.device, non_blocking=True)
        code = {torch.float64: _cabi.TVDK_F64, torch.float32: _cabi.TVDK_F32}[src.dtype]
        g = self.owned_geom()
        dst = self.field(name, comp)[GL * self.plane_elems:]
        with device_guard(self.device):
            rc = self.lib.pack(C.byref(g), code, src.data_ptr(), dst.data_ptr(), self.stream())
        self.lib.check(rc, "tvdk_pack")
        self._keep = src

    def get_owned(self, name: str, out_dtype=None, comp: int = 0) -> np.ndarray:
        n = self.nown[self.rank]
        out_dtype = self.cdtype if out_dtype is None else np.dtype(out_dtype)
        g = self.owned_geom()
        src = self.field(name, comp)[GL * self.plane_elems:]
        nvox = n * int(np.prod(self.shape[1:]))
        out = torch.empty(nvox, dtype=_TORCH[out_dtype], device=self.device)
        with device_guard(self.device):
            rc = self.lib.unpack(C.byref(g), _CODE[out_dtype], src.data_ptr(), out.data_ptr(), self.stream())
        self.lib.check(rc, "tvdk_unpack")
        if self.device.type == "cuda":
            host = torch.empty(out.shape, dtype=out.dtype, pin_memory=True)
            host.copy_(out, non_blocking=True)
            torch.cuda.current_stream(self.device).synchronize()
            res = host.numpy()
        else:
            res = out.numpy()
        return res.reshape((n,) + self.shape[1:])


class SlabGroup:
    """The slabs of one decomposition that THIS process drives (all of them, or one)."""

    def __init__(self, global_shape: Sequence[int], dim: int, dtype, mu, omega, ell, isotropic: bool, solver,
                 eps: Optional[float], local_ranks: Sequence[int], world: int, devices: Sequence, group=None):
        if len(global_shape) != dim:
            raise AssertionError("slab decomposition handles one volume (no batch axes)")
        if world > _cabi.TVDK_MAX_SLABS:
            raise ValueError(f"at most {_cabi.TVDK_MAX_SLABS} slabs")
        self.lib = lib = _cabi.get_library()
        self.spec = _solver_spec(solver)
        if self.spec.kind not in (_cabi.TVDK_SOLVER_JACOBI, _cabi.TVDK_SOLVER_CG):
            raise NotImplementedError("slab-decomposed solves support Jacobi and CG")
        if ell is None:
            ell = 2 * mu
        self.shape, self.dim, self.world, self.group = tuple(int(n) for n in global_shape), dim, world, group
        n0 = self.shape[0]
        if n0 < world:
            raise ValueError(f"cannot cut {n0} planes into {world} slabs")
        self.nown = [slab_range(n0, r, world)[1] - slab_range(n0, r, world)[0] for r in range(world)]
        if world > 1 and min(self.nown) < GH:
            raise ValueError("every slab needs at least 2 planes")
        self.cdtype = np.dtype(dtype)
        self.coefs = {"mu": mu, "omega": omega, "ell": ell}
        self.maps = [k for k, v in self.coefs.items() if isinstance(v, (list, tuple)) or np.ndim(v) != 0]
        for k, v in self.coefs.items():
            if k == "ell" and k not in self.maps:
                _check_positive("ell", v)
        self.seq = C.c_uint64(0)
        self.ranks: List[_Rank] = []
        for r, dev in zip(local_ranks, devices):
            dev = resolve_device(dev)
            rk = _Rank(lib, self.shape, r, world, self.nown, self.cdtype, dev, self.spec, isotropic, eps, self.maps)
            rk.comm.seq = C.pointer(self.seq)
            self.ranks.append(rk)
        self.eps = eps
        self._peer_maps = []  # (pointer, offset) of every arena opened over CUDA IPC
        self._closed = False
        self._map_arenas()
        for rk in self.ranks:  # scalar coefficients / pointers of the maps
            for k, v in self.coefs.items():
                c = getattr(rk.problem, k)
                if k in self.maps:
                    c.device_ptr, c.value, c.batch_stride = rk.base + rk.off[k], 0.0, 0
                else:
                    c.device_ptr, c.value, c.batch_stride = None, float(v), 0
        self._ready = False
        self.launched = 0

    # -- construction ------------------------------------------------------------------------------
    @classmethod
    def local(cls, global_shape, dim, dtype, mu, omega, ell, isotropic, solver, eps=None, world=2, devices=None):
        """All ``world`` slabs in this process: on ``devices[r]`` each, or all on the current device."""
        if devices is None:
            devices = [None] * world
        return cls(global_shape, dim, dtype, mu, omega, ell, isotropic, solver, eps, list(range(world)), world, devices)

    @classmethod
    def distributed(cls, global_shape, dim, dtype, mu, omega, ell, isotropic, solver, eps=None, group=None, device=None):
        """One slab per process of ``group`` (``torch.distributed``), this process drives its own."""
        import torch.distributed as dist

        rank = dist.get_rank(group) if dist.is_initialized() else 0
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        return cls(global_shape, dim, dtype, mu, omega, ell, isotropic, solver, eps, [rank], world, [device], group=group)

    def _map_arenas(self) -> None:
        """Fill ``comm.arena[r]`` of every local rank with rank r's arena as seen from that rank's device."""
        lib = self.lib
        local = {rk.rank: rk for rk in self.ranks}
        if len(local) == self.world:  # one process: plain pointers (+ peer access between distinct GPUs)
            for rk in self.ranks:
                for other in self.ranks:
                    rk.comm.arena[other.rank] = other.base
                    if rk.device != other.device and rk.device.type == "cuda":
                        with device_guard(rk.device):
                            lib.check(lib.enable_peer(other.device.index), "tvdk_enable_peer")
            return
        import torch.distributed as dist

        assert len(self.ranks) == 1, "either all slabs or exactly one per process"
        rk = self.ranks[0]
        handle = (C.c_ubyte * 64)()
        off = C.c_int64(0)
        with device_guard(rk.device):
            lib.check(lib.ipc_export(rk.base, C.byref(handle), C.byref(off)), "tvdk_ipc_export")
        mine = (bytes(handle), int(off.value), int(rk.arena_bytes))
        everyone = [None] * self.world
        dist.all_gather_object(everyone, mine, group=self.group)
        for r, (h, o, nbytes) in enumerate(everyone):
            if nbytes != rk.arena_bytes:
                raise RuntimeError("ranks disagree on the arena layout")
            if r == rk.rank:
                rk.comm.arena[r] = rk.base
                continue
            hb = (C.c_ubyte * 64).from_buffer_copy(h)
            ptr = C.c_void_p(0)
            with device_guard(rk.device):
                lib.check(lib.ipc_open(C.byref(hb), o, C.byref(ptr)), "tvdk_ipc_open")
            rk.comm.arena[r] = ptr.value
            self._peer_maps.append((ptr.value, o))
        dist.barrier(group=self.group)  # every header is zero and mapped before the first flag is written

    def close(self) -> None:
        """Unmap the peers' arenas and release this process' own (collective for ``distributed`` groups: every
        rank calls it).  An arena exported over CUDA IPC stays allocated on ITS device until every importer has
        closed the mapping, so dropping the last Python reference alone does not give the memory back."""
        if self._closed:
            return
        self._closed = True
        self._ready = False
        for rk in self.ranks:
            if rk.device.type == "cuda":
                torch.cuda.synchronize(rk.device)
        if self._peer_maps:
            import torch.distributed as dist

            dist.barrier(group=self.group)  # nobody still pushes into an arena that is about to be unmapped
            rk = self.ranks[0]
            with device_guard(rk.device):
                for ptr, o in self._peer_maps:
                    self.lib.check(self.lib.ipc_close(ptr, o), "tvdk_ipc_close")
            self._peer_maps = []
            dist.barrier(group=self.group)
        for rk in self.ranks:
            rk.arena = None
            rk._keep = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False

    # -- data ----------------------------------------------------------------------------------------
    def _arrays(self):
        n = len(self.ranks)
        P = (C.POINTER(_cabi.Problem) * n)(*[C.pointer(rk.problem) for rk in self.ranks])
        Cm = (C.POINTER(_cabi.SlabComm) * n)(*[C.pointer(rk.comm) for rk in self.ranks])
        S = (C.c_void_p * n)(*[rk.stream() for rk in self.ranks])
        devs = [(rk.device.index if rk.device.type == "cuda" else -1) for rk in self.ranks]
        Dv = (C.c_int * n)(*devs)
        return n, P, Cm, S, Dv

    def _exchange(self, names_comps, n_up: int, n_down: int) -> None:
        n, _, Cm, S, Dv = self._arrays()
        G = (C.POINTER(_cabi.Geom) * n)(*[C.pointer(rk.geom) for rk in self.ranks])
        k = len(names_comps)
        ptr_rows = []
        for rk in self.ranks:
            ptr_rows.append((C.c_void_p * k)(*[rk.base + rk.off[nm] + cp * rk.field_bytes for nm, cp in names_comps]))
        F = (C.POINTER(C.c_void_p) * n)(*[C.cast(row, C.POINTER(C.c_void_p)) for row in ptr_rows])
        rc = self.lib.slab_exchange(n, G, Cm, k, F, n_up, n_down, S, Dv)
        self.lib.check(rc, "tvdk_slab_exchange")

    def setup(self, owned_images: Sequence) -> None:
        """Upload every local rank's owned planes of the image (and of the per-voxel maps given at construction,
        as owned planes too), reset the state to u = img, t = 0."""
        assert len(owned_images) == len(self.ranks)
        for rk, img in zip(self.ranks, owned_images):
            rk.put_owned("f", img)
            for k in self.maps:
                v = self.coefs[k]
                if isinstance(v, (list, tuple)):  # one array of owned planes per slab
                    v = v[rk.rank]
                elif tuple(v.shape) == self.shape:  # the whole map: cut this rank's planes
                    v = v[rk.s:rk.s + rk.nown[rk.rank]]
                rk.put_owned(k, np.asarray(v, dtype=self.cdtype) if not isinstance(v, torch.Tensor) else v)
        if self.maps:
            self._exchange([(k, 0) for k in self.maps], GL, GH)
        self.reset()
        self._ready = True

    def reset(self) -> None:
        """u = img, t = 0 (device-side copies; the image stays resident)."""
        for rk in self.ranks:
            with device_guard(rk.device):
                rk.field("u").copy_(rk.field("f"))
                for d in range(self.dim):
                    rk.field("t", d).zero_()

    def solve(self, max_num_iter: int) -> int:
        """Enqueue the whole outer loop (``tvdk_solve_slab``); returns the number of outer iterations enqueued."""
        assert self._ready, "call setup() first"
        if max_num_iter > self.ranks[0].hist_cap:
            raise ValueError(f"at most {self.ranks[0].hist_cap} outer iterations per call")
        for rk in self.ranks:
            rk.problem.max_num_iter = int(max_num_iter)
        n, P, Cm, S, Dv = self._arrays()
        launched = C.c_int(0)
        rc = self.lib.solve_slab(n, P, Cm, S, Dv, C.byref(launched))
        self.lib.check(rc, "tvdk_solve_slab")
        self.launched = int(launched.value)
        return self.launched

    def check(self) -> None:
        """Raise if a flag wait timed out inside the kernels (a peer died or fell out of step)."""
        for rk in self.ranks:
            seq = C.c_int64(0)
            with device_guard(rk.device):
                self.lib.check(self.lib.slab_error(rk.base, rk.stream(), C.byref(seq)), "tvdk_slab_error")
            if seq.value:
                raise RuntimeError(f"slab rank {rk.rank}: timed out waiting for the peers' flags at step {seq.value}")

    def stats(self, i: int = 0) -> "_cabi.Stat":
        rk = self.ranks[i]
        o = rk.off["stats"]
        raw = rk.arena[o:o + C.sizeof(_cabi.Stat)].cpu().numpy().tobytes()
        return _cabi.Stat.from_buffer_copy(raw)

    def history(self, i: int = 0) -> List[float]:
        rk = self.ranks[i]
        n = self.stats(i).iters
        o = rk.off["hist"]
        return rk.arena[o:o + 8 * n].view(torch.float64).cpu().tolist()

    def results(self, out_dtype=None) -> List[np.ndarray]:
        """Owned planes of the result of every local rank."""
        return [rk.get_owned("u", out_dtype) for rk in self.ranks]


def split_bregman_tvd_slab_p2p(img_owned, global_shape, mu=1.0, omega=1.0, ell=None, dim: int = 3, max_num_iter: int = 100,
                               eps: Optional[float] = None, isotropic: bool = False, solver=None, group=None, device=None,
                               precision: Optional[str] = None, return_info: bool = False):
    """``split_bregman_tvd`` for a volume distributed over the ranks of ``group`` (one process per GPU), through
    ``tvdk_solve_slab``.  Arguments as ``darsia_b200.split_bregman_tvd_slab``."""
    if not isinstance(img_owned, np.ndarray) or img_owned.dtype not in (np.float32, np.float64):
        raise TypeError("float32/float64 ndarray expected")
    cdtype = img_owned.dtype if precision is None else np.dtype({"fp64": np.float64, "fp32": np.float32}[precision])
    with SlabGroup.distributed(global_shape, dim, cdtype, mu, omega, ell, isotropic, solver, eps=eps, group=group, device=device) as grp:
        grp.setup([img_owned])
        grp.solve(max_num_iter)
        out = grp.results(img_owned.dtype)[0]
        grp.check()
        hist = grp.history() if return_info else None
    return (out, hist) if return_info else out
