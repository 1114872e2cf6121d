"""Build the CPU thread-stepper of the kernels (TEST INFRASTRUCTURE ONLY; see tvdk_emu.cpp)."""

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SO = os.path.join(HERE, "libtvdk_emu.so")
SRCS = [
    os.path.join(HERE, "tvdk_emu.cpp"),
    os.path.join(ROOT, "darsia_b200", "csrc", "tvdk_core.h"),
    os.path.join(ROOT, "darsia_b200", "csrc", "tvdk_host.inl"),
    os.path.join(ROOT, "darsia_b200", "csrc", "tvdk_slab.inl"),
    os.path.join(ROOT, "include", "tvdk.h"),
]


def build(force: bool = False) -> str:
    newest = max(os.path.getmtime(s) for s in SRCS)
    if force or not os.path.exists(SO) or os.path.getmtime(SO) < newest:
        cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-o", SO, SRCS[0]]
        subprocess.run(cmd, check=True, cwd=HERE)
    return SO


if __name__ == "__main__":
    print(build(force=True))
