"""CPU, build container only: the NumPy oracle against the VERBATIM reference on seeded random problems.

Runs only where ``/root/reference`` exists (it does not exist on the GPU box, and nothing marked ``gpu`` ever
reads it).  The committed fixtures (tests/golden/*.npz) pin the oracle on 42 chosen cases; this test widens the
pin to random shapes / coefficients / solvers, restricted to what the reference itself can run (SURVEY Q3/Q5/Q6:
no trailing batch axes; isotropic only for 2-D fp64 with scalar mu and ell).  The reference is executed in a
subprocess because loading it installs stub ``darsia`` / ``skimage`` modules and the SciPy ``tol`` shim."""

import os
import subprocess
import sys

import pytest

from conftest import ROOT

REF = os.environ.get("DARSIA_REFERENCE", "/root/reference/src/darsia")

SCRIPT = r'''
import sys
import numpy as np
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[1] + "/tests"); sys.path.insert(0, sys.argv[1] + "/tests/golden")
import generate_golden as gg
import oracle
ref = gg.load_reference()
rng = np.random.default_rng(2024)
worst = 0.0
n = 0
while n < 30:
    dim = int(rng.choice([2, 3]))
    shape = (int(rng.integers(3, 30)), int(rng.integers(3, 40))) if dim == 2 else tuple(int(v) for v in rng.integers(3, 14, 3))
    dt = np.float64 if rng.random() < 0.8 else np.float32
    kind = str(rng.choice(["j", "jtol", "cg", "cgtol", "mg"]))
    het = bool(rng.random() < 0.4) and kind != "mg"
    iso = bool(rng.random() < 0.4) and dim == 2 and dt == np.float64
    if kind == "mg" and min(shape) < 4:
        continue
    if kind in ("cg", "cgtol") and dt == np.float32:
        continue  # SciPy promotes internally; the fp32 CG route of the reference is not a stable target
    img = gg.phantom(shape, int(rng.integers(0, 10**6)), dt)
    kw = dict(mu=0.1, omega=float(rng.choice([0.5, 1.0, 2.0])), ell=float(rng.choice([0.5, 1.0])))
    if het:
        kw["omega"] = gg.coeff_field(shape, 7, 0.5, 2.0).astype(dt)
        if not iso:
            kw["mu"] = gg.coeff_field(shape, 6, 0.05, 0.1, 0.1).astype(dt)
            if rng.random() < 0.5:  # per-voxel ell: Jacobi multiplies it after, CG puts it inside the divergence (Q18)
                kw["ell"] = gg.coeff_field(shape, 5, 0.5, 1.5).astype(dt)
    K = int(rng.integers(1, 9))
    def mk(api):
        return {"j": lambda: api.Jacobi(maxiter=K), "jtol": lambda: api.Jacobi(maxiter=20, tol=1e-2),
                "cg": lambda: api.CG(maxiter=min(K, 6), tol=0.0), "cgtol": lambda: api.CG(maxiter=30, tol=1e-3),
                "mg": lambda: api.MG(depth=0, smoother_iterations=min(K, 4), maxiter=2)}[kind]()
    common = dict(dim=dim, max_num_iter=int(rng.integers(1, 7)), isotropic=iso, eps=None if rng.random() < 0.8 else 5e-3)
    a = ref.split_bregman_tvd(img.copy(), solver=mk(ref), **common, **kw)
    b = oracle.split_bregman_tvd(img.copy(), solver=mk(oracle), **common, **kw)
    err = float(np.linalg.norm(a.astype(np.float64) - b.astype(np.float64)) / np.linalg.norm(a.astype(np.float64)))
    tol = 1e-13 if dt == np.float64 else 1e-5
    assert a.shape == b.shape and a.dtype == b.dtype and err <= tol, (shape, dt, kind, het, iso, common, err)
    worst = max(worst, err if dt == np.float64 else 0.0)
    n += 1
print("ok", n, "worst fp64", worst)
'''


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference checkout is only present in the build container")
def test_oracle_matches_verbatim_reference_on_random_problems():
    res = subprocess.run([sys.executable, "-c", SCRIPT, ROOT], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    assert res.stdout.strip().splitlines()[-1].startswith("ok 30")


SCRIPT_MG = r'''
import sys
import numpy as np
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[1] + "/tests"); sys.path.insert(0, sys.argv[1] + "/tests/golden")
import generate_golden as gg
import oracle
ref = gg.load_reference()
rng = np.random.default_rng(77)
worst = 0.0
for n in range(16):
    dim = int(rng.choice([2, 3]))
    depth = int(rng.integers(0, 3))
    lo = 2 ** (depth + 1)
    shape = tuple(int(v) for v in rng.integers(lo, 40 if dim == 2 else 15, dim))
    img = gg.phantom(shape, int(rng.integers(0, 10**6)))
    which = str(rng.choice(["both", "omega", "ell"]))
    om = gg.coeff_field(shape, 7 + n, 0.5, 2.0) if which in ("both", "omega") else float(rng.choice([0.5, 1.0, 2.0]))
    el = gg.coeff_field(shape, 5 + n, 0.5, 1.5) if which in ("both", "ell") else float(rng.choice([0.5, 1.0]))
    cp = lambda c: c.copy() if isinstance(c, np.ndarray) else c
    tol = None if rng.random() < 0.7 else 1e-3
    def mk(api):
        # (heterogeneous is decided by the constructor's arguments, mg.py:70-73)
        return api.MG(depth=depth, smoother_iterations=int(rng_k), maxiter=int(rng_m), tol=tol, dim=dim,
                      mass_coeff=cp(om) if isinstance(om, np.ndarray) else None,
                      diffusion_coeff=cp(el) if isinstance(el, np.ndarray) else None)
    rng_k, rng_m = rng.integers(1, 5), rng.integers(1, 4)
    common = dict(dim=dim, max_num_iter=int(rng.integers(1, 5)), isotropic=bool(rng.random() < 0.3) and dim == 2 and not isinstance(el, np.ndarray),
                  mu=0.1, omega=cp(om), ell=cp(el))
    a = ref.split_bregman_tvd(img.copy(), solver=mk(ref), **common)
    common["omega"], common["ell"] = cp(om), cp(el)
    b = oracle.split_bregman_tvd(img.copy(), solver=mk(oracle), **common)
    err = float(np.linalg.norm(a - b) / np.linalg.norm(a))
    assert a.shape == b.shape and err <= 1e-13, (shape, depth, which, tol, common["max_num_iter"], err)
    worst = max(worst, err)
print("ok 16 worst", worst)
'''


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference checkout is only present in the build container")
def test_oracle_matches_verbatim_reference_mg_with_per_voxel_coefficients():
    """MG constructed with per-voxel mass / diffusion coefficients (mg.py:70-73, 134-149, 170-192): random shapes (odd
    extents on every level), depths 0-2, either or both coefficients as arrays, with and without the MG tolerance."""
    res = subprocess.run([sys.executable, "-c", SCRIPT_MG, ROOT], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    assert res.stdout.strip().splitlines()[-1].startswith("ok 16")
