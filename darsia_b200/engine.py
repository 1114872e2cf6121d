"""Device-resident engine behind ``split_bregman_tvd``: geometry, field allocation, layout
conversion and the call into ``tvdk_solve``.

``Engine`` owns every device buffer (as torch tensors) and hands raw pointers to the C ABI.
It is also what ``bench.py`` times: ``upload`` -> ``solve`` -> ``download``.
"""

from __future__ import annotations

import ctypes as C
import os
import threading
from dataclasses import dataclass
from typing import Optional, Sequence, Tuple, Union

import numpy as np
import torch

from . import _cabi
from ._runtime import device_guard, resolve_device, stream_ptr

_TORCH_DTYPE = {_cabi.TVDK_F64: torch.float64, _cabi.TVDK_F32: torch.float32}
_NP_CODE = {np.dtype(np.float64): _cabi.TVDK_F64, np.dtype(np.float32): _cabi.TVDK_F32}


@dataclass
class SolverSpec:
    """What the kernels need to know about a ``darsia.Solver`` descriptor."""

    kind: int  # _cabi.TVDK_SOLVER_*
    maxiter: int
    tol: Optional[float]


@dataclass
class SolveInfo:
    """Per-problem outcome of one ``Engine.solve``."""

    iters: np.ndarray  # outer iterations performed per problem
    inc: np.ndarray  # last relative increment per problem
    done: np.ndarray  # eps criterion fired
    history: Optional[np.ndarray]  # (max_num_iter, nb) relative increments (NaN where not run)
    energy: Optional[np.ndarray]  # (max_num_iter, nb) energies when verbose
    launched: int  # outer iterations enqueued by the host
    cg_pred_dev: Optional[np.ndarray] = None  # single-pass CG: largest relative deviation of the predicted rho
    nonfinite: Optional[np.ndarray] = None  # a relative increment came out NaN / Inf (the problem was frozen there)


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


class Engine:
    """One batch of ``nb`` independent ``dim``-dimensional problems of the same spatial shape."""

    def __init__(
        self,
        spatial_shape: Sequence[int],
        nb: int = 1,
        dtype=np.float64,
        device=None,
        tile: Tuple[int, int, int] = (0, 0, 0),
        slab: Optional[Tuple[int, int, int, int]] = None,
    ) -> None:
        """``spatial_shape`` is the LOCAL extent.  ``slab = (g0_off, g0_n, z_begin, z_end)`` marks
        the array as one slab of a volume cut along axis 0 (see ``tvdk_geom_t``); by default the
        array is the whole volume."""
        self.lib = _cabi.get_library()
        self.device = resolve_device(device)
        self.spatial = tuple(int(n) for n in spatial_shape)
        self.dim = len(self.spatial)
        if self.dim not in (2, 3):
            raise ValueError(f"dim must be 2 or 3, got {self.dim}")
        self.nb = int(nb)
        self.np_dtype = np.dtype(dtype)
        if self.np_dtype not in _NP_CODE:
            raise TypeError(f"compute dtype must be float32 or float64, got {self.np_dtype}")
        self.code = _NP_CODE[self.np_dtype]
        self.tdtype = _TORCH_DTYPE[self.code]
        self.tile = tuple(int(x) for x in tile)
        g = _cabi.Geom()
        g.dim = self.dim
        g.dtype = self.code
        for a in range(3):
            g.n[a] = self.spatial[a] if a < self.dim else 1
        g.pitch = self.lib.pitch(self.code, self.spatial[-1])
        g.nb = self.nb
        if slab is None:
            slab = (0, self.spatial[0], 0, self.spatial[0])
        g.g0_off, g.g0_n, g.z_begin, g.z_end = (int(v) for v in slab)
        self.slab = tuple(int(v) for v in slab)
        self.geom = g
        self.field_elems = int(self.lib.field_elems(C.byref(g)))  # one problem
        if self.field_elems <= 0:
            raise ValueError(self.lib.last_error().decode())
        self.batch_elems = self.field_elems * self.nb  # one field over all problems
        self.nvox = int(np.prod(self.spatial))
        self.plane_elems = self.field_elems // self.spatial[0]  # one plane of axis 0 (contiguous)
        self._op_ws: Optional[torch.Tensor] = None

    # ------------------------------------------------------------------ allocation / layout
    def stream(self) -> int:
        return stream_ptr(self.lib, self.device)

    def empty_fields(self, n: int = 1) -> torch.Tensor:
        """``n`` fields (each over all problems), uninitialised."""
        return torch.empty(n * self.batch_elems, dtype=self.tdtype, device=self.device)

    def zero_fields(self, n: int = 1) -> torch.Tensor:
        return torch.zeros(n * self.batch_elems, dtype=self.tdtype, device=self.device)

    def _geom_nb(self, nb: int) -> _cabi.Geom:
        g = _cabi.Geom.from_buffer_copy(self.geom)
        g.nb = nb
        return g

    def to_device(self, arr) -> torch.Tensor:
        """Host array -> flat device tensor in host order (a plain H2D copy); a tensor that already
        lives on the device is used as it is (device-resident callers: no copy)."""
        if isinstance(arr, torch.Tensor):
            return arr.to(self.device, non_blocking=True).contiguous().reshape(-1)
        t = torch.from_numpy(np.ascontiguousarray(arr).reshape(-1))
        if self.device.type != "cuda" or t.is_pinned() or t.numel() * t.element_size() < (32 << 20):
            return t.to(self.device, non_blocking=True)
        return self._stage_pageable(t)

    _STAGE: dict = {}  # (device index, dtype, chunk) -> free pairs of pinned chunks (a call borrows one pair)
    _STAGE_LOCK = threading.Lock()
    _STAGE_POOL = None
    # host threads copying into the pinned chunks / chunk size (tools/h2d_probe.py measures both on a given box)
    # (B200 box, 672 MB fp64: 1 thread 12 GB/s, 4 threads 34 GB/s, pinned source 55 GB/s -- profiles/r2_h2d_probe.txt;
    #  the default leaves two cores per rank of a one-process-per-GPU job)
    STAGE_THREADS = int(os.environ.get("DARSIA_B200_STAGE_THREADS", "0")) or max(
        1, min(4, (os.cpu_count() or 1) // (2 * max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1"))))))
    STAGE_CHUNK_MB = int(os.environ.get("DARSIA_B200_STAGE_CHUNK_MB", "16"))

    def _stage_pageable(self, t: torch.Tensor) -> torch.Tensor:
        """A large PAGEABLE host array -> device through two pinned chunks: the host memcpy of chunk i + 1 overlaps the
        DMA of chunk i, and ``STAGE_THREADS`` host threads share each memcpy (a plain ``cudaMemcpy`` from pageable memory
        stages internally on one thread: 11 GB/s on the 672 MB image of BASELINE configs[1])."""
        chunk = (Engine.STAGE_CHUNK_MB << 20) // t.element_size()
        nt = max(1, Engine.STAGE_THREADS)
        key = (self.device.index, t.dtype, chunk)
        with Engine._STAGE_LOCK:  # (concurrent callers -- tvd_series' two host threads -- each borrow their own pair)
            free = Engine._STAGE.setdefault(key, [])
            bufs = free.pop() if free else None
            if nt > 1 and Engine._STAGE_POOL is None:
                from concurrent.futures import ThreadPoolExecutor

                Engine._STAGE_POOL = ThreadPoolExecutor(nt)
        if bufs is None:
            bufs = [torch.empty(chunk, dtype=t.dtype, pin_memory=True) for _ in range(2)]
        out = torch.empty(t.numel(), dtype=t.dtype, device=self.device)
        src = t.numpy()
        evs = [None, None]
        stream = torch.cuda.current_stream(self.device)
        for i, off in enumerate(range(0, t.numel(), chunk)):
            m = min(chunk, t.numel() - off)
            b = i & 1
            if evs[b] is not None:
                evs[b].synchronize()
            dst = bufs[b].numpy()
            if nt == 1:
                dst[:m] = src[off:off + m]
            else:  # (NumPy releases the GIL while it copies)
                step = (m + nt - 1) // nt
                list(Engine._STAGE_POOL.map(
                    lambda k: np.copyto(dst[k * step:min(m, (k + 1) * step)], src[off + k * step:off + min(m, (k + 1) * step)]),
                    range(nt)))
            out[off:off + m].copy_(bufs[b][:m], non_blocking=True)
            evs[b] = torch.cuda.Event()
            evs[b].record(stream)
        for ev in evs:  # the staging chunks are shared: do not hand them to the next caller while a DMA still reads them
            if ev is not None:
                ev.synchronize()
        with Engine._STAGE_LOCK:
            Engine._STAGE[key].append(bufs)
        return out

    def pack(self, inter: torch.Tensor, nb: Optional[int] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Host-order device tensor ``(voxels, nb)`` -> planar padded field(s) (``tvdk_pack``)."""
        nb = self.nb if nb is None else nb
        if inter.numel() != self.nvox * nb:
            raise AssertionError(f"expected {self.nvox * nb} elements, got {inter.numel()}")
        src_code = {torch.float64: _cabi.TVDK_F64, torch.float32: _cabi.TVDK_F32}[inter.dtype]
        if out is None:
            # zero-filled so that the pad columns hold finite values
            out = torch.zeros(self.field_elems * nb, dtype=self.tdtype, device=self.device)
        g = self._geom_nb(nb)
        with device_guard(self.device):
            rc = self.lib.pack(C.byref(g), src_code, inter.data_ptr(), out.data_ptr(), self.stream())
        self.lib.check(rc, "tvdk_pack")
        return out

    def unpack(self, field: torch.Tensor, dst_dtype=None, nb: Optional[int] = None) -> torch.Tensor:
        """Planar field(s) -> host-order device tensor ``(voxels, nb)`` (``tvdk_unpack``)."""
        nb = self.nb if nb is None else nb
        dst_dtype = self.np_dtype if dst_dtype is None else np.dtype(dst_dtype)
        code = _NP_CODE[dst_dtype]
        out = torch.empty(self.nvox * nb, dtype=_TORCH_DTYPE[code], device=self.device)
        g = self._geom_nb(nb)
        with device_guard(self.device):
            rc = self.lib.unpack(C.byref(g), code, field.data_ptr(), out.data_ptr(), self.stream())
        self.lib.check(rc, "tvdk_unpack")
        return out

    def upload(self, img) -> torch.Tensor:
        """Host (or device) image ``(*spatial, *batch)`` -> field (H2D copy + layout kernel)."""
        return self.pack(self.to_device(img))

    def expand_state(self, t: torch.Tensor, mu, ell, isotropic: bool, full_shape=None):
        """Compressed state ``t`` -> the reference's ``(d, b)`` (``split_bregman_tvd.py:100-101``) as
        planar fields (``dim`` components each), via ``tvdk_expand_state``."""
        d = self.empty_fields(self.dim)
        b = self.empty_fields(self.dim)
        muc, k1 = self.coef(mu, full_shape, "mu")
        elc, k2 = self.coef(ell, full_shape, "ell")
        with device_guard(self.device):
            rc = self.lib.expand_state(C.byref(self.geom), 1 if isotropic else 0, C.byref(muc), C.byref(elc),
                                       t.data_ptr(), self.batch_elems, d.data_ptr(), b.data_ptr(), self.batch_elems,
                                       self.stream())
        self.lib.check(rc, "tvdk_expand_state")
        return d, b

    def components_to_host_order(self, planar: torch.Tensor, dst_dtype, shape) -> torch.Tensor:
        """``dim`` planar component fields -> device tensor ``(*shape, dim)`` (component axis last, as
        the reference keeps d and b)."""
        comps = [self.unpack(planar[a * self.batch_elems:(a + 1) * self.batch_elems], dst_dtype).reshape(shape)
                 for a in range(self.dim)]
        return torch.stack(comps, dim=-1)

    def download(self, field: torch.Tensor, out_dtype=None, shape=None) -> np.ndarray:
        """Field -> host array ``(*spatial, *batch)``."""
        inter = self.unpack(field, out_dtype)
        if self.device.type == "cuda":
            # D2H into pinned memory (torch's caching host allocator reuses the block)
            host_t = torch.empty(inter.shape, dtype=inter.dtype, pin_memory=True)
            host_t.copy_(inter, non_blocking=True)
            torch.cuda.current_stream(self.device).synchronize()
            host = host_t.numpy()
        else:
            host = inter.numpy()
        return host.reshape(shape if shape is not None else (*self.spatial, self.nb))

    def coef(self, value, full_shape: Optional[tuple] = None, name: str = "coefficient"):
        """Scalar or per-voxel coefficient -> (``tvdk_coef_t``, keep-alive tensor)."""
        c = _cabi.Coef()
        if isinstance(value, torch.Tensor) and value.device == self.device and value.dim() == 1:
            # already a packed field (shared map or per problem)
            if value.numel() not in (self.field_elems, self.batch_elems):
                raise AssertionError(f"{name}: packed field has the wrong size")
            c.device_ptr = value.data_ptr()
            c.value = 0.0
            c.batch_stride = self.field_elems if (self.nb > 1 and value.numel() == self.batch_elems) else 0
            return c, value
        if isinstance(value, torch.Tensor):  # an N-d map, possibly already on the device
            arr = value if value.dim() > 0 else np.asarray(float(value))
        else:
            arr = np.asarray(value)
        if arr.ndim == 0:
            c.device_ptr = None
            c.value = float(arr)
            c.batch_stride = 0
            return c, None
        if tuple(arr.shape) == self.spatial:
            nbc = 1
        elif full_shape is not None and tuple(arr.shape) == tuple(full_shape):
            nbc = self.nb
        else:
            raise AssertionError(
                f"{name} of shape {arr.shape} matches neither the spatial shape {self.spatial} nor the image shape"
            )
        if isinstance(arr, torch.Tensor):
            field = self.pack(self.to_device(arr.to(self.tdtype)), nb=nbc)
        else:
            field = self.pack(self.to_device(arr.astype(self.np_dtype, copy=False)), nb=nbc)
        c.device_ptr = field.data_ptr()
        c.value = 0.0
        c.batch_stride = self.field_elems if nbc > 1 else 0
        return c, field

    def op_workspace(self) -> torch.Tensor:
        if self._op_ws is None:
            n = int(self.lib.op_workspace_bytes(C.byref(self.geom)))
            self._op_ws = torch.empty(n, dtype=torch.uint8, device=self.device)
        return self._op_ws

    # ------------------------------------------------------------------ the outer loop
    def solve(
        self,
        f: torch.Tensor,
        u: torch.Tensor,
        t: torch.Tensor,
        mu,
        omega,
        ell,
        isotropic: bool,
        solver: SolverSpec,
        max_num_iter: int,
        eps: Optional[float],
        x0_db: Optional[Tuple[torch.Tensor, torch.Tensor]] = None,
        verbose_every: int = 0,
        record_history: bool = False,
        fused: bool = True,
        check_every: int = 8,
        full_shape: Optional[tuple] = None,
        sync: bool = True,
        adaptive_mask: Optional[np.ndarray] = None,
    ) -> Optional[SolveInfo]:
        """Run ``tvdk_solve`` on resident fields: ``u`` (in: initial iterate, out: result) and the
        compressed state ``t`` are updated in place.  With ``sync=False`` nothing is read back
        (``None`` is returned) and the call only enqueues work.
        """
        lib = self.lib
        p = _cabi.Problem()
        p.geom = self.geom
        p.isotropic = 1 if isotropic else 0
        p.solver_kind = solver.kind
        p.solver_maxiter = int(solver.maxiter)
        p.solver_tol = -1.0 if solver.tol is None else float(solver.tol)
        p.max_num_iter = int(max_num_iter)
        p.eps = -1.0 if eps is None else float(eps)
        keep = []
        for name, val in (("mu", mu), ("omega", omega), ("ell", ell)):
            c, k = self.coef(val, full_shape, name)
            setattr(p, name, c)
            keep.append(k)
        p.f, p.u, p.t = f.data_ptr(), u.data_ptr(), t.data_ptr()
        p.t_comp_stride = self.batch_elems
        if x0_db is not None:
            p.x0_d, p.x0_b = x0_db[0].data_ptr(), x0_db[1].data_ptr()
            p.x0_comp_stride = self.batch_elems
        p.verbose_every = int(verbose_every)
        p.check_every = int(check_every)
        p.fused = int(fused)  # False/True, or 2: single-pass CG iterations in 3-D as well
        p.tile_x, p.tile_y, p.zchunk = self.tile
        if adaptive_mask is not None:
            adaptive_mask = np.ascontiguousarray(adaptive_mask, dtype=np.uint8)
            if adaptive_mask.shape != (int(max_num_iter),):
                raise AssertionError("adaptive_mask needs one flag per outer iteration")
            p.adaptive_mask = adaptive_mask.ctypes.data  # host memory, read during the call only
        nbytes = int(lib.workspace_bytes(C.byref(p)))
        if nbytes == 0:
            raise ValueError(lib.last_error().decode())
        ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        stats = torch.zeros(self.nb * C.sizeof(_cabi.Stat), dtype=torch.uint8, device=self.device)
        hist = ehist = None
        if record_history or verbose_every > 0:
            hist = torch.full((max(max_num_iter, 1) * self.nb,), float("nan"), dtype=torch.float64, device=self.device)
        if verbose_every > 0:
            ehist = torch.full((max(max_num_iter, 1) * self.nb,), float("nan"), dtype=torch.float64, device=self.device)
        p.workspace, p.workspace_bytes = ws.data_ptr(), nbytes
        p.stats = stats.data_ptr()
        p.history, p.energy_history = _ptr(hist), _ptr(ehist)
        launched = C.c_int(0)
        with device_guard(self.device):
            rc = lib.solve(C.byref(p), self.stream(), C.byref(launched))
        lib.check(rc, "tvdk_solve")
        self._last = (ws, stats, hist, ehist, keep)  # keep alive until the stream has consumed them
        if not sync:
            return None
        raw = stats.cpu().numpy().tobytes()  # synchronises the stream
        recs = (_cabi.Stat * self.nb).from_buffer_copy(raw)
        return SolveInfo(
            iters=np.array([r.iters for r in recs]),
            inc=np.array([r.inc for r in recs]),
            done=np.array([bool(r.done) for r in recs]),
            history=None if hist is None else hist.cpu().numpy().reshape(-1, self.nb),
            energy=None if ehist is None else ehist.cpu().numpy().reshape(-1, self.nb),
            launched=int(launched.value),
            cg_pred_dev=np.array([r.cg_pred_dev for r in recs]),
            nonfinite=np.array([bool(r.nonfinite) for r in recs]),
        )
