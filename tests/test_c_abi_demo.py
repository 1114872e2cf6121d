"""The C ABI from plain C (tests/c_abi/tvdk_c_demo.c): no Python, no PyTorch on the call path.

CPU: include/tvdk.h is valid C99 and libtvdk.so resolves every symbol the demo uses (compile + link only).
GPU: the demo runs one split_bregman_tvd call per solver family on the B200 and matches the oracle."""

import os
import subprocess
import sys

import numpy as np
import pytest

import oracle
from conftest import ROOT, rel_l2

sys.path.insert(0, os.path.join(ROOT, "tests", "c_abi"))
import build_demo  # noqa: E402

needs_lib = pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "darsia_b200", "csrc", "libtvdk.so")),
                               reason="libtvdk.so not built (run __graft_entry__.build())")


@needs_lib
def test_demo_compiles_and_links_as_c99(tmp_path):
    exe = build_demo.build(str(tmp_path))
    assert os.access(exe, os.X_OK)


@pytest.mark.gpu
@pytest.mark.parametrize("iso,solver,maxiter,outer", [(0, 0, 1, 12), (1, 1, 6, 5), (0, 0, 7, 4)])
def test_demo_matches_oracle_on_gpu(tmp_path, iso, solver, maxiter, outer):
    from golden.generate_golden import phantom

    exe = build_demo.build(str(tmp_path))
    img = phantom((150, 301), 21)
    fin, fout = tmp_path / "in.bin", tmp_path / "out.bin"
    img.tofile(fin)
    res = subprocess.run([exe, str(fin), str(fout), "150", "301", str(iso), str(solver), str(maxiter), str(outer)],
                         capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    assert f"iterations {outer}" in res.stdout
    out = np.fromfile(fout, dtype=np.float64).reshape(img.shape)
    sol = oracle.CG(maxiter=maxiter, tol=0.0) if solver else oracle.Jacobi(maxiter=maxiter)
    ref = oracle.split_bregman_tvd(img, mu=0.1, omega=1.0, ell=1.0, dim=2, isotropic=bool(iso), max_num_iter=outer,
                                   solver=sol)
    assert rel_l2(out, ref) <= 1e-12
