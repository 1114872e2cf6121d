"""pytest configuration: the ``gpu`` marker and shared helpers."""

import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def golden_case_names():
    names = sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.endswith(".npz") and f not in ("operators.npz", "h1.npz"))
    return names


def load_golden(name):
    """Return (arrays dict, params dict) of one golden case."""
    with np.load(os.path.join(GOLDEN_DIR, name + ".npz")) as z:
        arrays = {k: z[k] for k in z.files if k != "params"}
        params = json.loads(str(z["params"]))
    return arrays, params


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / (den if den > 0 else 1.0))


def run_case_with(api, arrays, params):
    """Run one golden case through an API namespace exposing the reference's names.

    ``api`` is either the ``oracle`` package or ``darsia_b200`` (same signatures as
    ``darsia``): ``tvd``, ``split_bregman_tvd``, ``Jacobi``, ``CG``.
    """
    p = dict(params)
    entry = p.pop("entry")
    has_x0 = p.pop("x0")
    spec = dict(p.pop("solver"))
    kind = spec.pop("kind")
    if spec.pop("hetero", False):  # MG decides `heterogeneous` from its constructor arguments (mg.py:70-73)
        if "coef_omega" in arrays:
            spec["mass_coeff"] = arrays["coef_omega"].copy()
        if "coef_ell" in arrays:
            spec["diffusion_coeff"] = arrays["coef_ell"].copy()
    solver = {"jacobi": api.Jacobi, "cg": api.CG, "mg": getattr(api, "MG", None)}[kind](**spec)
    kwargs = dict(p)
    kwargs.pop("parity_tol", None)
    kwargs["solver"] = solver
    every = kwargs.pop("adaptive_every", None)
    if every:  # same schedule as tests/golden/generate_golden.py:run_case
        kwargs["adaptive"] = lambda it, every=every: it % every == every - 1
    img = arrays["img"].copy()
    if has_x0:
        kwargs["x0"] = (arrays["x0_u"].copy(), arrays["x0_d"].copy(), arrays["x0_b"].copy())
    coefs = {k[5:]: v for k, v in arrays.items() if k.startswith("coef_")}
    if entry == "tvd":
        if "mu" in coefs:
            kwargs["weight"] = coefs.pop("mu")
        kwargs.update(coefs)
        return api.tvd(img, **kwargs)
    kwargs.update(coefs)
    return api.split_bregman_tvd(img, **kwargs)


def tolerance_for(dtype, params=None):
    """Parity bar of BASELINE.json north_star: 1e-10 (fp64), 1e-5 (fp32) relative L2.  The few
    ill-conditioned fixtures (adaptive ell) carry their own, justified in generate_golden.py."""
    if params and params.get("parity_tol"):
        return float(params["parity_tol"])
    return 1e-10 if np.dtype(dtype) == np.float64 else 1e-5


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN_DIR
