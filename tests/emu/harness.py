"""Inject the CPU kernel emulator into darsia_b200 -- for the CPU test-suite ONLY.

``darsia_b200`` itself only ever loads ``darsia_b200/csrc/libtvdk.so`` and requires a CUDA
device.  The tests that must run without a GPU (host logic, index/halo/boundary logic of the
kernels, multi-rank gloo paths) replace the process-wide library handle by the emulator built
from the same kernel source (tests/emu/tvdk_emu.cpp), in which "device pointers" are host
pointers and torch tensors live on the CPU.
"""

import contextlib
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
if HERE not in sys.path:
    sys.path.insert(0, HERE)


def emulator_library():
    import build_emu
    from darsia_b200 import _cabi

    return _cabi.Library(build_emu.build())


@contextlib.contextmanager
def emulated():
    from darsia_b200 import _cabi

    saved = _cabi._LIB
    _cabi._LIB = emulator_library()
    try:
        yield _cabi._LIB
    finally:
        _cabi._LIB = saved
