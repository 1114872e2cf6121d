"""A time series of optical images through the streaming pipeline -- the loop of ``manager/analysisbase.py:205-227``
(``restoration(img)`` once per image of the series, with a mask-scaled per-pixel weight as the TOML workflow builds it,
``presets/workflows/restoration.py:108-148``) with the frames sharded over the GPUs of the node.

    python examples/series_stream.py [frames] [H] [W]                                     # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        examples/series_stream.py 64 2000 3500                                            # frame-sharded, no collective

Frames are produced one at a time (here synthesised; in DarSIA they come from ``darsia.imread`` of the next file) and
handed back one at a time: the series never exists as one array on the host.  H2D of sub-batch k+1, the solve of
sub-batch k and D2H of sub-batch k-1 overlap (``darsia_b200/series.py``).
"""

import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))  # run from a checkout
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import darsia_b200 as darsia  # noqa: E402

n_frames = int(sys.argv[1]) if len(sys.argv) > 1 else 16
shape = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (1000, 1750)

rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local)
if world > 1:  # rendezvous only: every frame is an independent problem, there is no data-path collective
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

# per-pixel weight = mu scaled by a mask (small outside the region of interest) and a per-pixel fidelity weight
yy, xx = np.indices(shape, sparse=True)
mask = ((yy > shape[0] // 10) & (xx > shape[1] // 10)).astype(np.float64)
weight = 0.1 * (0.05 + 0.95 * mask)  # (kept positive: ell defaults to 2 * weight, split_bregman_tvd.py:62-63)
omega = 0.5 + 1.5 * (xx / shape[1]) * np.ones(shape)


def frame(k: int) -> np.ndarray:
    """Frame k of the series (a moving front + noise); stands for darsia.imread(paths[k]).img."""
    rng = np.random.default_rng(k)
    front = (xx + 0 * yy) < (0.2 + 0.6 * k / max(1, n_frames - 1)) * shape[1]
    return np.clip(0.2 + 0.6 * front + 0.1 * rng.standard_normal(shape), 0.0, 1.0)


received = {}


def sink(k: int, denoised: np.ndarray) -> None:
    """Called in frame order with a view of a pinned staging buffer: reduce it (or copy it) before returning."""
    received[k] = (float(denoised.mean()), float(np.abs(np.diff(denoised, axis=1)).mean()))


t0 = time.perf_counter()
_, (s, e) = darsia.tvd_series_stream(
    frame, n_frames=n_frames, dim=2, sink=sink, rank=rank, world=world,
    mu=weight, omega=omega, ell=None, isotropic=False, solver=darsia.Jacobi(maxiter=5), max_num_iter=100,
)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
noisy_tv = float(np.abs(np.diff(frame(s), axis=1)).mean())
print(f"rank {rank}/{world}: frames [{s}, {e}) of {n_frames} at {shape[0]} x {shape[1]} in {dt:.2f} s "
      f"({(e - s) * shape[0] * shape[1] * 100 / dt / 1e6:.0f} Mvoxel*iter/s incl. frame synthesis and copies); "
      f"mean |d/dx|: {noisy_tv:.4f} noisy -> {received[s][1]:.4f} denoised")
if world > 1:
    dist.barrier(device_ids=[local])
    dist.destroy_process_group()
