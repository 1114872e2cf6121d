#!/usr/bin/env python
"""A few small solves covering every kernel family (march kernels 2-D / 3-D, the TMA-staged kernels, the slab driver's
push / finalize kernels in lock step) -- the program compute-sanitizer runs (tools/gpu_r2_e.sh):
    compute-sanitizer --tool racecheck|initcheck python tools/sanitize_small.py
Results are checked against the oracle so that the run also proves the instrumented kernels still compute the right thing."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import darsia_b200 as da  # noqa: E402
import oracle  # noqa: E402
from darsia_b200.slab_p2p import SlabGroup  # noqa: E402
from golden.generate_golden import coeff_field, phantom  # noqa: E402


def check(name, out, ref, tol=1e-10):
    err = float(np.linalg.norm(out - ref) / np.linalg.norm(ref))
    print(f"{name}: rel L2 = {err:.2e}")
    assert err <= tol, name


img = phantom((40, 70), 1)
vol = phantom((12, 20, 40), 2)
kw2 = dict(mu=0.1, omega=1.0, ell=1.0, dim=2, max_num_iter=3)
kw3 = dict(mu=0.1, omega=1.0, ell=1.0, dim=3, max_num_iter=3)
for iso in (False, True):
    check(f"2-D J1 iso={iso}", da.split_bregman_tvd(img, isotropic=iso, solver=da.Jacobi(maxiter=1), **kw2),
          oracle.split_bregman_tvd(img, isotropic=iso, solver=oracle.Jacobi(maxiter=1), **kw2))
    check(f"3-D J1 (TMA) iso={iso}", da.split_bregman_tvd(vol, isotropic=iso, solver=da.Jacobi(maxiter=1), **kw3),
          oracle.split_bregman_tvd(vol, isotropic=iso, solver=oracle.Jacobi(maxiter=1), **kw3))
check("2-D CG(4) single pass", da.split_bregman_tvd(img, isotropic=True, solver=da.CG(maxiter=4, tol=0.0), **kw2),
      oracle.split_bregman_tvd(img, isotropic=True, solver=oracle.CG(maxiter=4, tol=0.0), **kw2))
check("3-D CG(3)", da.split_bregman_tvd(vol, isotropic=True, solver=da.CG(maxiter=3, tol=0.0), **kw3),
      oracle.split_bregman_tvd(vol, isotropic=True, solver=oracle.CG(maxiter=3, tol=0.0), **kw3))
check("2-D J(5) blocked sweeps", da.split_bregman_tvd(img, solver=da.Jacobi(maxiter=5), **kw2),
      oracle.split_bregman_tvd(img, solver=oracle.Jacobi(maxiter=5), **kw2))
mu = 0.1 * coeff_field((40, 70), 5, 0.5, 1.0, 0.1)
check("2-D J1 per-pixel mu", da.split_bregman_tvd(img, mu=mu, omega=1.0, ell=1.0, dim=2, max_num_iter=3, solver=da.Jacobi(maxiter=1)),
      oracle.split_bregman_tvd(img, mu=mu, omega=1.0, ell=1.0, dim=2, max_num_iter=3, solver=oracle.Jacobi(maxiter=1)))
for solver in ("j1", "cg"):
    mk = (lambda m: m.Jacobi(maxiter=1)) if solver == "j1" else (lambda m: m.CG(maxiter=3, tol=0.0))
    grp = SlabGroup.local(vol.shape, 3, vol.dtype, 0.1, 1.0, 1.0, solver == "cg", mk(da), world=2)
    grp.setup([np.ascontiguousarray(vol[slice(*da.slab_range(vol.shape[0], r, 2))]) for r in range(2)])
    grp.solve(3)
    grp.check()
    check(f"slab driver, 2 slabs in lock step, {solver}", np.concatenate(grp.results(), axis=0),
          oracle.split_bregman_tvd(vol, isotropic=solver == "cg", solver=mk(oracle), **kw3))
print("sanitize_small ok")
