"""Frame-sharded TV denoising of image series (BASELINE config 3): no data-path collective.

A DarSIA series keeps time as the axis right after the spatial ones (``image/image.py:72-79``),
so once trailing axes are independent problems every frame -- and every colour
channel of every frame -- is its own TVD problem.  Frames are dealt to the ranks of the
process group in contiguous blocks; each rank solves its frames in sub-batches that fit its
HBM; nothing is exchanged until the (optional) final gather of the results.
"""

from __future__ import annotations

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
    starts = list(range(0, e - s, batch))

    def solve(b0: int) -> None:
        sel = (slice(None),) * dim + (slice(b0, min(b0 + batch, e - s)),)
        out[sel] = split_bregman_tvd(np.ascontiguousarray(mine[sel]), dim=dim, **kwargs)

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
