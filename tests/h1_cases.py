"""H1 regularisation golden cases (reference outputs in tests/golden/h1.npz)."""
import os

import numpy as np

from conftest import GOLDEN_DIR

CASES = {
    "vol_jac_tol": ("vol", dict(mu=1, omega=1, dim=3), ("jacobi", dict(maxiter=30, tol=1e-6))),
    "vol_cg": ("vol", dict(mu=0.5, omega=2.0, dim=3), ("cg", dict(maxiter=8, tol=0.0))),
    "rgb_jac5": ("rgb", dict(mu=0.7, omega=1.0, dim=2), ("jacobi", dict(maxiter=5))),
    "rgb_jac_tol": ("rgb", dict(mu=0.7, omega=1.0, dim=2), ("jacobi", dict(maxiter=40, tol=1e-3))),
    "vol_mg": ("vol", dict(mu=1.0, omega=1.0, dim=3), ("mg", dict(depth=1, smoother_iterations=3, maxiter=4))),
    "vol_mg_hetero": ("vol", dict(mu="vol_mu", omega="vol_om", dim=3),
                      ("mg", dict(depth=1, smoother_iterations=3, maxiter=3, dim=3, mass_coeff="vol_om", diffusion_coeff="vol_mu"))),
}


def load():
    with np.load(os.path.join(GOLDEN_DIR, "h1.npz")) as z:
        return {k: z[k] for k in z.files}


def run(api, data, name):
    src, kw, (kind, skw) = CASES[name]
    # (strings name arrays stored next to the images: per-voxel coefficients)
    kw = {k: (data[v].copy() if isinstance(v, str) else v) for k, v in kw.items()}
    skw = {k: (data[v].copy() if isinstance(v, str) else v) for k, v in skw.items()}
    solver = {"jacobi": api.Jacobi, "cg": api.CG, "mg": api.MG}[kind](**skw)
    return api.H1_regularization(data[src].copy(), solver=solver, **kw), data[name]
