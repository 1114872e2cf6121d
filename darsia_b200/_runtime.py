"""Device plumbing: which device tensors live on and which stream kernels are enqueued on.

PyTorch is used for device memory, streams and host<->device copies only.
"""

from __future__ import annotations

from typing import Optional, Union

import torch

from ._cabi import Library, get_library


_CHECKED: set = set()  # device indices that passed tvdk_device_ok


def resolve_device(device: Optional[Union[str, int, torch.device]] = None) -> torch.device:
    """Return the CUDA device to run on; raise if there is none (no CPU fallback).

    The only way to get a CPU device out of this function is the kernel *emulator* that the
    CPU test-suite injects (``tests/emu``); the shipped ``libtvdk.so`` requires sm_100.
    """
    lib = get_library()
    if lib.is_emulator:
        return torch.device("cpu")
    if not torch.cuda.is_available():
        raise RuntimeError(
            "darsia_b200 needs a CUDA device (B200, sm_100a); none is visible and there is no CPU fallback"
        )
    if device is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    elif isinstance(device, int):
        dev = torch.device("cuda", device)
    else:
        dev = torch.device(device)
        if dev.type != "cuda":
            raise ValueError(f"device must be a CUDA device, got {dev}")
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
    if dev.index not in _CHECKED:  # once per device and process
        with torch.cuda.device(dev):
            if not lib.device_ok():
                raise RuntimeError(lib.last_error().decode())
        _CHECKED.add(dev.index)
    return dev


def stream_ptr(lib: Library, dev: torch.device) -> int:
    """``cudaStream_t`` of torch's current stream on ``dev`` (0 for the emulator)."""
    if lib.is_emulator:
        return 0
    return int(torch.cuda.current_stream(dev).cuda_stream)


class device_guard:
    """``with device_guard(dev):`` makes ``dev`` current for raw CUDA calls inside libtvdk."""

    def __init__(self, dev: torch.device) -> None:
        self.ctx = torch.cuda.device(dev) if dev.type == "cuda" else None

    def __enter__(self):
        if self.ctx is not None:
            self.ctx.__enter__()
        return self

    def __exit__(self, *exc):
        if self.ctx is not None:
            return self.ctx.__exit__(*exc)
        return False
