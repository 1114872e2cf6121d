"""Build tests/c_abi/tvdk_c_demo.c (plain C against include/tvdk.h + libtvdk.so + libcudart)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CUDA = os.environ.get("CUDA_HOME", "/usr/local/cuda")
CSRC = os.path.join(ROOT, "darsia_b200", "csrc")


def build(out_dir: str) -> str:
    exe = os.path.join(out_dir, "tvdk_c_demo")
    cmd = ["gcc", "-std=c99", "-O1", "-Wall", "-Werror", "-o", exe, os.path.join(HERE, "tvdk_c_demo.c"),
           "-I", os.path.join(ROOT, "include"), "-I", os.path.join(CUDA, "include"),
           "-L", CSRC, "-ltvdk", "-L", os.path.join(CUDA, "lib64"), "-lcudart", f"-Wl,-rpath,{CSRC}",
           f"-Wl,-rpath,{os.path.join(CUDA, 'lib64')}"]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    return exe
