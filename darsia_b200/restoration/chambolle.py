"""Chambolle TV denoising on the GPU -- the algorithm behind ``darsia.tvd(img)``'s DEFAULT method.

The reference forwards ``method="chambolle"`` to ``skimage.restoration.denoise_tv_chambolle(img,
weight=..., max_num_iter=..., eps=...)`` (``restoration/tvd.py:77-83``): every axis of the array is
spatial (no ``channel_axis``).  scikit-image is a third-party dependency that is not installed in
the build image, so this path is a restatement of its published algorithm
(``oracle/chambolle_oracle.py``) and its parity is UNPINNED -- it is provided so that the default
``tvd()`` call has a GPU path at all; the contract of this package is the ``"heterogeneous bregman"``
method.  One libtvdk kernel per iteration (``tvdk_chambolle``), device-resident loop, the stopping
rule ``|E_prev - E| < eps * E_0`` evaluated on the device.
"""

from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from .. import _cabi
from .._runtime import device_guard
from ..engine import Engine


def denoise_tv_chambolle(image, weight: float = 0.1, eps: float = 2.0e-4, max_num_iter: int = 200, *, device=None,
                         return_info: bool = False):
    """``skimage.restoration.denoise_tv_chambolle`` (all axes spatial) for 2-D and 3-D float arrays or CUDA
    tensors; float32 stays float32.  Returns the denoised array (and, with ``return_info``, the number of
    iterations that ran and the energies)."""
    on_device = isinstance(image, torch.Tensor)
    if not on_device and not isinstance(image, np.ndarray):
        raise TypeError(f"Input type {type(image)} not supported.")
    dt = {torch.float32: np.float32, torch.float64: np.float64}.get(image.dtype) if on_device else image.dtype
    if dt not in (np.float32, np.float64):
        raise TypeError(f"float32/float64 images only, got {image.dtype} (scikit-image would rescale with img_as_float)")
    if image.ndim not in (2, 3):
        raise NotImplementedError("chambolle on the GPU: 2-D and 3-D arrays (every axis is spatial, as in scikit-image)")
    if not weight > 0:
        raise ValueError("weight must be positive")
    max_num_iter = int(max_num_iter)
    if max_num_iter < 1:
        raise ValueError("max_num_iter must be >= 1")
    shape = tuple(image.shape)
    if on_device and device is None and image.is_cuda:
        device = image.device
    eng = Engine(shape, nb=1, dtype=np.dtype(dt), device=device)
    lib, dim = eng.lib, eng.dim
    f = eng.upload(image)
    out = eng.empty_fields(1)
    p0, p1 = eng.zero_fields(dim), eng.empty_fields(dim)
    ws = eng.op_workspace()
    stats = torch.zeros(C.sizeof(_cabi.Stat), dtype=torch.uint8, device=eng.device)
    hist = torch.full((max_num_iter,), float("nan"), dtype=torch.float64, device=eng.device) if return_info else None
    c = _cabi.Chambolle()
    c.geom = eng.geom
    c.weight, c.eps, c.max_num_iter, c.check_every = float(weight), float(eps), max_num_iter, 8
    c.f, c.out, c.p0, c.p1 = f.data_ptr(), out.data_ptr(), p0.data_ptr(), p1.data_ptr()
    c.p_comp_stride = eng.batch_elems
    c.workspace, c.workspace_bytes = ws.data_ptr(), ws.numel()
    c.stats = stats.data_ptr()
    c.history = None if hist is None else hist.data_ptr()
    launched = C.c_int(0)
    with device_guard(eng.device):
        rc = lib.chambolle(C.byref(c), eng.stream(), C.byref(launched))
    lib.check(rc, "tvdk_chambolle")
    res = eng.unpack(out, dt).reshape(shape) if on_device else eng.download(out, dt, shape)
    if not return_info:
        return res
    rec = _cabi.Stat.from_buffer_copy(stats.cpu().numpy().tobytes())
    energies = hist.cpu().numpy()[: rec.iters]
    return res, dict(iterations=int(rec.iters), energies=[float(e) for e in energies], launched=int(launched.value))
