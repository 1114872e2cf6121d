"""Frame-sharded TV denoising of image series (BASELINE config 3): no data-path collective.

A DarSIA series keeps time as the axis right after the spatial ones (``image/image.py:72-79``),
so once trailing axes are independent problems every frame -- and every colour
channel of every frame -- is its own TVD problem.  Frames are dealt to the ranks of the
process group in contiguous blocks; each rank solves its frames in sub-batches that fit its
HBM; nothing is exchanged until the (optional) final gather of the results.
"""

from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np

from .restoration.split_bregman_tvd import split_bregman_tvd


def frame_range(n_frames: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block of frames [s, e) owned by ``rank`` (as even as possible)."""
    base, rem = divmod(n_frames, world)
    s = rank * base + min(rank, rem)
    return s, s + base + (1 if rank < rem else 0)


def frames_per_batch(frame_voxels: int, itemsize: int, dim: int, budget_bytes: float = 40e9) -> int:
    """How many frames to solve at once: about 2*dim + 6 resident fields per frame (image, two
    iterates, two compressed states, solver work, staging)."""
    per_frame = frame_voxels * itemsize * (2 * dim + 6)
    return max(1, int(budget_bytes // max(per_frame, 1)))


def _run_overlapped(solve, starts, device) -> None:
    """Two worker threads, one CUDA stream each: copies of one sub-batch overlap kernels of the other."""
    import threading

    import torch

    from ._runtime import resolve_device

    dev = resolve_device(device)
    errors = []

    def worker(items) -> None:
        try:
            if dev.type == "cuda":
                torch.cuda.set_device(dev)
                with torch.cuda.stream(torch.cuda.Stream(dev)):
                    for b0 in items:
                        solve(b0)
            else:  # the CPU test-suite's kernel emulator
                for b0 in items:
                    solve(b0)
        except BaseException as exc:  # re-raised in the caller's thread
            errors.append(exc)

    threads = [threading.Thread(target=worker, args=(starts[i::2],)) for i in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]


def tvd_series(
    series: np.ndarray,
    dim: int = 2,
    rank: Optional[int] = None,
    world: Optional[int] = None,
    batch: Optional[int] = None,
    gather: bool = False,
    group=None,
    overlap: bool = True,
    **kwargs,
):
    """Denoise the frames ``series[..., t]`` owned by this rank.

    Args:
        series: array ``(*spatial, T)`` or ``(*spatial, T, C)`` -- every rank passes the same array
            (or at least its own frames; only frames ``frame_range(T, rank, world)`` are read).
        dim: spatial dimension.
        rank, world: default to the ``torch.distributed`` default group (or a single rank).
        batch: frames per device-resident sub-batch (default: sized from the HBM budget).
        gather: if True the full denoised series is assembled on every rank (``all_gather_object``
            of the host arrays -- convenience for modest sizes, not a data-path collective).
        overlap: process the sub-batches from two host threads, each on its own CUDA stream, so that the
            host<->device copies of one sub-batch overlap the kernels of the other (libtvdk keeps no global
            state; results are identical -- every sub-batch is an independent solve).
        **kwargs: forwarded to ``split_bregman_tvd`` (``mu``, ``omega``, ``ell``, ``solver``, ...).
    Returns ``(frames, (s, e))``: the denoised frames of this rank and their index range.
    """
    import torch.distributed as dist

    if rank is None or world is None:
        if dist.is_available() and dist.is_initialized():
            rank, world = dist.get_rank(group), dist.get_world_size(group)
        else:
            rank, world = 0, 1
    if series.ndim < dim + 1:
        raise AssertionError("series needs a time axis after the spatial axes")
    T = series.shape[dim]
    s, e = frame_range(T, rank, world)
    mine = series[(slice(None),) * dim + (slice(s, e),)]
    out = np.empty_like(mine)
    if batch is None:
        # (two sub-batches are resident at a time when the copies are overlapped)
        batch = frames_per_batch(int(np.prod(series.shape[:dim])) * int(np.prod(series.shape[dim + 1:], dtype=np.int64)),
                                 series.dtype.itemsize, dim)
    # (gridDim.z carries the problem index: at most 65535 problems -- frames x channels -- per launch)
    chans = int(np.prod(series.shape[dim + 1:], dtype=np.int64))
    batch = max(1, min(int(batch), 65535 // max(1, chans)))
    starts = list(range(0, e - s, batch))
    full_shape = tuple(series.shape)

    def solve(b0: int) -> None:
        sel = (slice(None),) * dim + (slice(b0, min(b0 + batch, e - s)),)
        kw = dict(kwargs)
        for name in ("mu", "omega", "ell"):  # maps given for the WHOLE series: this sub-batch's frames of them
            c = kw.get(name)
            if c is not None and np.ndim(c) != 0 and tuple(np.shape(c)) == full_shape:
                kw[name] = np.ascontiguousarray(c[(slice(None),) * dim + (slice(s + b0, s + min(b0 + batch, e - s)),)])
        out[sel] = split_bregman_tvd(np.ascontiguousarray(mine[sel]), dim=dim, **kw)

    if overlap and len(starts) > 1:
        _run_overlapped(solve, starts, kwargs.get("device"))
    else:
        for b0 in starts:
            solve(b0)
    if gather and world > 1:
        parts = [None] * world
        dist.all_gather_object(parts, (s, e, out), group=group)
        full = np.empty_like(series)
        for ps, pe, arr in parts:
            full[(slice(None),) * dim + (slice(ps, pe),)] = arr
        return full, (s, e)
    return out, (s, e)


# ---------------------------------------------------------------------------------------------------------
# Streaming series (BASELINE config 3 as specified: 512 frames of 4000 x 7000 do not fit one host array)
# ---------------------------------------------------------------------------------------------------------
def tvd_series_stream(
    frames,
    n_frames: Optional[int] = None,
    dim: int = 2,
    sink=None,
    batch: Optional[int] = None,
    rank: Optional[int] = None,
    world: Optional[int] = None,
    group=None,
    device=None,
    out_dtype=None,
    **kwargs,
):
    """Denoise a series frame by frame WITHOUT materialising it: the loop ``manager/analysisbase.py:205-227`` runs
    over the images of a series, with ``restoration(img)`` (``tvd.py:45-63``) replaced by a three-stage pipeline:

        H2D of sub-batch k+1 (pinned staging ring, copy stream)  |  solve of sub-batch k (``tvdk_solve``)  |
        D2H of sub-batch k-1 (pinned ring, copy-out stream) + delivery to ``sink``

    so that the copies hide behind the kernels.  Every frame (and every colour channel of it) is an independent
    problem; frames are dealt to ranks in contiguous blocks -- no data-path collective.

    Args:
        frames: either a callable ``k -> ndarray`` (frame ``k`` of the series, ``(*spatial)`` or ``(*spatial, C)``;
            needs ``n_frames``; only this rank's frames ``frame_range(n_frames, rank, world)`` are requested), or an
            iterable over THIS rank's frames in order.
        sink: callable ``(k, frame)`` receiving each denoised frame (global index ``k``); ``frame`` is a view of a
            pinned staging buffer that is recycled two sub-batches later -- copy it if it must live longer.  With
            ``sink=None`` the frames are copied into a list that is returned.
        batch: frames per sub-batch (default: sized so that three sub-batches fit a 40 GB budget, at most 8).
        kwargs: as ``split_bregman_tvd`` (``mu``, ``omega``, ``ell``, ``isotropic``, ``solver``, ``max_num_iter``, ``eps``,
            ``precision``, ``fused``); per-voxel ``mu`` / ``omega`` / ``ell`` maps have the spatial shape and are shared by
            all frames (uploaded once).
    Returns ``(frames_or_None, (s, e))``.
    """
    import threading
    import queue

    import torch
    import torch.distributed as dist

    from . import _cabi
    from ._runtime import device_guard, resolve_device
    from .engine import Engine
    from .restoration.split_bregman_tvd import _check_positive, _solver_spec

    if rank is None or world is None:
        if dist.is_available() and dist.is_initialized():
            rank, world = dist.get_rank(group), dist.get_world_size(group)
        else:
            rank, world = 0, 1
    if callable(frames):
        if n_frames is None:
            raise ValueError("a frame callable needs n_frames")
        s, e = frame_range(int(n_frames), rank, world)
        source = (frames(k) for k in range(s, e))
    else:
        source = iter(frames)
        s, e = 0, None
    # ---- arguments of the solve (same checks as split_bregman_tvd) ----
    mu, omega, ell = kwargs.pop("mu", 1.0), kwargs.pop("omega", 1.0), kwargs.pop("ell", None)
    isotropic = bool(kwargs.pop("isotropic", False))
    spec = _solver_spec(kwargs.pop("solver", None))
    if spec.kind not in (_cabi.TVDK_SOLVER_JACOBI, _cabi.TVDK_SOLVER_CG):
        raise NotImplementedError("the streaming pipeline supports Jacobi and CG")
    max_num_iter = int(kwargs.pop("max_num_iter", 100))
    eps = kwargs.pop("eps", None)
    precision = kwargs.pop("precision", None)
    fused = kwargs.pop("fused", True)
    if kwargs:
        raise TypeError(f"unexpected arguments {sorted(kwargs)}")
    if ell is None:
        ell = 2 * mu
    _check_positive("ell", ell)

    first = next(source, None)
    if first is None:
        return ([] if sink is None else None), (s, s if e is None else e)
    first = np.asarray(first)
    if first.dtype not in (np.float32, np.float64):
        raise TypeError(f"float32/float64 frames only, got {first.dtype}")
    spatial = tuple(first.shape[:dim])
    chans = int(np.prod(first.shape[dim:], dtype=np.int64))
    cdtype = first.dtype if precision is None else np.dtype({"fp64": np.float64, "fp32": np.float32, "float64": np.float64,
                                                             "float32": np.float32}[precision])
    odtype = first.dtype if out_dtype is None else np.dtype(out_dtype)
    fvox = int(np.prod(spatial)) * chans
    if batch is None:
        batch = max(1, min(8, frames_per_batch(fvox, cdtype.itemsize, dim, budget_bytes=40e9 / 3)))
    dev = resolve_device(device)
    cuda = dev.type == "cuda"
    B = int(batch)
    eng = Engine(spatial, nb=B * chans, dtype=cdtype, device=dev)
    eng1 = Engine(spatial, nb=chans, dtype=cdtype, device=dev)  # layout kernels work frame by frame
    coefs = []
    for name, val in (("mu", mu), ("omega", omega), ("ell", ell)):
        if np.ndim(val) != 0:
            if tuple(np.shape(val)) != spatial:
                raise AssertionError(f"{name}: per-voxel maps of a streamed series have the spatial shape {spatial}")
            val = eng.coef(np.asarray(val), name=name)[1]  # packed once, shared by all frames
        coefs.append(val)
    tt = {np.dtype(np.float64): torch.float64, np.dtype(np.float32): torch.float32}
    NSLOT = 3
    pin = dict(pin_memory=True) if cuda else {}
    h_in = [torch.empty((B, fvox), dtype=tt[np.dtype(first.dtype)], **pin) for _ in range(NSLOT)]
    h_out = [torch.empty((B, fvox), dtype=tt[odtype], **pin) for _ in range(NSLOT)]
    d_in = [torch.empty((B, fvox), dtype=tt[np.dtype(first.dtype)], device=dev) for _ in range(2)]
    d_out = [torch.empty((B, fvox), dtype=tt[odtype], device=dev) for _ in range(2)]
    fld = [(eng.empty_fields(1), eng.empty_fields(1), eng.empty_fields(dim)) for _ in range(2)]
    if cuda:
        st_in, st_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        st_cmp = torch.cuda.current_stream(dev)
    collected = [] if sink is None else None
    errors = []
    free_in: "queue.Queue" = queue.Queue()
    for i in range(NSLOT):
        free_in.put(i)
    ready: "queue.Queue" = queue.Queue(maxsize=NSLOT)
    done: "queue.Queue" = queue.Queue()
    free_out: "queue.Queue" = queue.Queue()
    for i in range(NSLOT):
        free_out.put(i)

    def loader():  # host frames -> pinned staging (plain memcpy, releases the GIL); frames that ARE pinned go as they are
        try:
            k, cur, n, slot = s, first, 0, free_in.get()
            direct = [None] * B
            while cur is not None:
                a = np.asarray(cur)
                if a.shape[:dim] != spatial or a.size != fvox:
                    raise AssertionError(f"frame {k}: shape {a.shape} differs from the first frame's")
                src = None
                if cuda and a.dtype == first.dtype and a.flags.c_contiguous:
                    tsrc = torch.from_numpy(a.reshape(-1))
                    if tsrc.is_pinned():  # page-locked by the caller: DMA straight from it, no staging copy
                        src = tsrc
                if src is None:
                    h_in[slot][n].numpy()[...] = a.reshape(-1)
                direct[n] = src
                n += 1
                k += 1
                cur = next(source, None)
                if n == B or cur is None:
                    ready.put((slot, n, k - n, direct[:n]))
                    if cur is not None:
                        n, slot = 0, free_in.get()
                        direct = [None] * B
            ready.put(None)
        except BaseException as exc:  # noqa: BLE001
            errors.append(exc)
            ready.put(None)

    def deliver():  # pinned results -> sink
        try:
            while True:
                item = done.get()
                if item is None:
                    return
                oslot, n, k0, ev = item
                if ev is not None:
                    ev.synchronize()
                for i in range(n):
                    fr = h_out[oslot][i].numpy().reshape(spatial + tuple(first.shape[dim:]))
                    if sink is None:
                        collected.append(fr.copy())
                    else:
                        sink(k0 + i, fr)
                free_out.put(oslot)
        except BaseException as exc:  # noqa: BLE001
            errors.append(exc)

    tl, td = threading.Thread(target=loader, daemon=True), threading.Thread(target=deliver, daemon=True)
    tl.start()
    td.start()
    in_code = {np.dtype(np.float64): _cabi.TVDK_F64, np.dtype(np.float32): _cabi.TVDK_F32}
    it = 0
    ev_free = [None, None]  # the compute stream is done reading d_in[i]
    ev_dout = [None, None]  # the copy-out stream is done reading d_out[i]
    total = 0
    while True:
        item = ready.get()
        if item is None:
            break
        slot, n, k0, direct = item
        di = it & 1
        f, u, t = fld[di]
        e_cur = eng if n == B else Engine(spatial, nb=n * chans, dtype=cdtype, device=dev)
        nfe = e_cur.batch_elems
        with device_guard(dev):
            if cuda:
                with torch.cuda.stream(st_in):
                    if ev_free[di] is not None:
                        st_in.wait_event(ev_free[di])
                    if all(t is None for t in direct):
                        d_in[di][:n].copy_(h_in[slot][:n], non_blocking=True)
                    else:
                        for i in range(n):
                            d_in[di][i].copy_(h_in[slot][i] if direct[i] is None else direct[i], non_blocking=True)
                    ev_in = torch.cuda.Event()
                    ev_in.record(st_in)
                st_cmp.wait_event(ev_in)
            else:
                d_in[di][:n].copy_(h_in[slot][:n])
            for i in range(n):  # frame i, channels interleaved -> problems i * chans .. of the planar batch
                rc = eng1.lib.pack(C.byref(eng1.geom), in_code[np.dtype(first.dtype)], d_in[di][i].data_ptr(),
                                   f[i * chans * eng.field_elems:].data_ptr(), eng1.stream())
                eng1.lib.check(rc, "tvdk_pack")
            if cuda:
                ev_free[di] = torch.cuda.Event()  # the layout kernels have consumed d_in[di]
                ev_free[di].record(st_cmp)
                ev_in.synchronize()  # the pinned staging slot may be refilled once the H2D copy has run
            free_in.put(slot)
            u[:nfe].copy_(f[:nfe])
            tview = t[:dim * nfe]
            tview.zero_()
            e_cur.solve(f[:nfe], u[:nfe], tview, coefs[0], coefs[1], coefs[2], isotropic, spec, max_num_iter, eps, fused=fused,
                        sync=False)
            if cuda and ev_dout[di] is not None:
                st_cmp.wait_event(ev_dout[di])  # the D2H copy of sub-batch k - 2 has drained d_out[di]
            for i in range(n):
                rc = eng1.lib.unpack(C.byref(eng1.geom), in_code[odtype], u[i * chans * eng.field_elems:].data_ptr(),
                                     d_out[di][i].data_ptr(), eng1.stream())
                eng1.lib.check(rc, "tvdk_unpack")
            oslot = free_out.get()
            if cuda:
                ev_cmp = torch.cuda.Event()
                ev_cmp.record(st_cmp)
                with torch.cuda.stream(st_out):
                    st_out.wait_event(ev_cmp)
                    h_out[oslot][:n].copy_(d_out[di][:n], non_blocking=True)
                    ev_out = torch.cuda.Event()
                    ev_out.record(st_out)
                ev_dout[di] = ev_out  # (d_out[di] is rewritten two sub-batches later: the wait sits there, not here)
                done.put((oslot, n, k0, ev_out))
            else:
                h_out[oslot][:n].copy_(d_out[di][:n])
                done.put((oslot, n, k0, None))
        total += n
        it += 1
        if errors:
            break
    done.put(None)
    td.join()
    tl.join(timeout=1.0)
    if errors:
        raise errors[0]
    if e is None:
        e = s + total
    return collected, (s, e)
