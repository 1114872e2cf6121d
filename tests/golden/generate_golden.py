"""Generate golden vectors for the TVD hot path from the UNMODIFIED DarSIA reference.

Run in the build container (where ``/root/reference`` exists)::

    python tests/golden/generate_golden.py

``import darsia`` fails there (matplotlib, scikit-image, ... are absent), so the
eight hot-path source files are executed VERBATIM, from where they lie under
``/root/reference/src/darsia``, inside a stub ``darsia`` namespace plus a stub
``skimage`` exposing the ``img_as_*`` identities the path uses on float arrays.
The only shims are ``scipy.sparse.linalg.cg(tol=...)`` -> ``rtol=...`` (SciPy >= 1.14
renamed the argument; same meaning) and ``np.lib.pad = np.pad`` (NumPy 2
removed the alias that ``utils/linear_solvers/mg.py:233`` still uses).  Nothing is copied into the repo;
only the inputs/outputs are stored, as ``tests/golden/<case>.npz``.

Each case stores its inputs (``img`` and any coefficient arrays), a JSON ``params``
string and the reference output ``out`` (plus intermediate operator outputs for the
operator-level cases).  A fresh solver object is used per call.
"""

from __future__ import annotations

import json
import os
import sys

import numpy as np

REF = os.environ.get("DARSIA_REFERENCE", "/root/reference/src/darsia")
OUT = os.path.dirname(os.path.abspath(__file__))


def load_reference():
    """Execute the reference hot-path files verbatim under a stub namespace (``oracle/ref_loader.py``)."""
    root = os.path.dirname(os.path.dirname(OUT))
    if root not in sys.path:
        sys.path.insert(0, root)
    from oracle.ref_loader import load_reference as _load

    return _load(REF if os.path.isdir(REF) else None)


def phantom(shape, seed, dtype=np.float64):
    """Piecewise-constant blocks + Gaussian noise, clipped to [0, 1]."""
    rng = np.random.default_rng(seed)
    idx = np.indices(shape)
    cells = [max(2, n // 4) for n in shape]
    checker = sum((idx[a] // cells[a]) for a in range(len(shape))) % 3
    base = np.choose(checker, [0.2, 0.5, 0.8])
    img = base + 0.1 * rng.standard_normal(shape)
    return np.clip(img, 0.0, 1.0).astype(dtype)


def coeff_field(shape, seed, lo, hi, zero_frac=0.0):
    rng = np.random.default_rng(seed)
    c = lo + (hi - lo) * rng.random(shape)
    if zero_frac > 0:
        c[rng.random(shape) < zero_frac] = 0.0
    return c


def make_solver(da, spec, arrays=None):
    kind = spec["kind"]
    kw = {k: v for k, v in spec.items() if k not in ("kind", "hetero")}
    if spec.get("hetero"):  # MG decides `heterogeneous` from its CONSTRUCTOR arguments (mg.py:70-73)
        if "omega" in arrays:
            kw["mass_coeff"] = arrays["omega"]
        if "ell" in arrays:
            kw["diffusion_coeff"] = arrays["ell"]
    return {"jacobi": da.Jacobi, "cg": da.CG, "mg": da.MG}[kind](**kw)


CASES = []


def case(name, shape, seed, dtype="float64", entry="direct", arrays=None, x0=False, **params):
    CASES.append(dict(name=name, shape=shape, seed=seed, dtype=dtype, entry=entry,
                      arrays=arrays or {}, x0=x0, params=params))


# --- solver-level end-to-end cases (fixed iteration counts) ---------------------------
J1 = {"kind": "jacobi", "maxiter": 1}
case("aniso2d_j1", (37, 52), 11, mu=0.1, omega=1.0, ell=1.0, dim=2, max_num_iter=25,
     isotropic=False, solver=J1)
case("aniso2d_j1_ellnone", (40, 33), 12, mu=0.15, omega=0.7, dim=2, max_num_iter=25,
     isotropic=False, solver=J1)
case("aniso2d_j20", (37, 52), 13, mu=0.1, omega=0.1, ell=1.0, dim=2, max_num_iter=12,
     isotropic=False, solver={"kind": "jacobi", "maxiter": 20})
case("aniso2d_j2", (18, 70), 14, mu=0.2, omega=1.0, ell=0.5, dim=2, max_num_iter=25,
     isotropic=False, solver={"kind": "jacobi", "maxiter": 2})
case("iso2d_j1", (37, 52), 15, mu=0.1, omega=1.0, ell=1.0, dim=2, max_num_iter=25,
     isotropic=True, solver=J1)
case("iso2d_j3", (41, 36), 16, mu=0.1, omega=1.0, ell=1.0, dim=2, max_num_iter=25,
     isotropic=True, solver={"kind": "jacobi", "maxiter": 3})
case("iso2d_cg7", (37, 52), 17, mu=0.1, omega=1.0, ell=1.0, dim=2, max_num_iter=25,
     isotropic=True, solver={"kind": "cg", "maxiter": 7, "tol": 0.0})
case("iso2d_cg10", (64, 48), 18, mu=0.1, omega=1.0, ell=1.0, dim=2, max_num_iter=20,
     isotropic=True, solver={"kind": "cg", "maxiter": 10, "tol": 0.0})
case("aniso2d_cg5", (37, 52), 19, mu=0.1, omega=1.0, ell=1.0, dim=2, max_num_iter=25,
     isotropic=False, solver={"kind": "cg", "maxiter": 5, "tol": 0.0})
case("aniso2d_cg_tol", (37, 52), 20, mu=0.1, omega=1.0, ell=1.0, dim=2, max_num_iter=15,
     isotropic=False, solver={"kind": "cg", "maxiter": 50, "tol": 1e-3})
case("aniso3d_j1", (13, 17, 11), 21, mu=0.1, omega=1.0, ell=1.0, dim=3, max_num_iter=25,
     isotropic=False, solver=J1)
case("aniso3d_j2", (9, 12, 20), 22, mu=0.1, omega=0.5, ell=1.0, dim=3, max_num_iter=20,
     isotropic=False, solver={"kind": "jacobi", "maxiter": 2})
case("aniso3d_cg6", (13, 17, 11), 23, mu=0.1, omega=1.0, ell=1.0, dim=3, max_num_iter=20,
     isotropic=False, solver={"kind": "cg", "maxiter": 6, "tol": 0.0})
# --- per-voxel coefficients ---------------------------------------------------------
case("aniso2d_j1_hetero", (37, 52), 24, arrays={"mu": (0.05, 0.1, 0.1), "omega": (0.5, 2.0, 0.0)},
     ell=1.0, dim=2, max_num_iter=25, isotropic=False, solver=J1)
case("aniso2d_cg5_hetero", (37, 52), 25, arrays={"mu": (0.05, 0.1, 0.1), "omega": (0.5, 2.0, 0.0)},
     ell=1.0, dim=2, max_num_iter=20, isotropic=False,
     solver={"kind": "cg", "maxiter": 5, "tol": 0.0})
case("aniso2d_j1_hetero_ell", (30, 41), 26, arrays={"ell": (0.5, 1.5, 0.0)}, mu=0.1, omega=1.0,
     dim=2, max_num_iter=20, isotropic=False, solver=J1)
case("aniso2d_cg5_hetero_ell", (30, 41), 27, arrays={"ell": (0.5, 1.5, 0.0)}, mu=0.1, omega=1.0,
     dim=2, max_num_iter=20, isotropic=False, solver={"kind": "cg", "maxiter": 5, "tol": 0.0})
case("aniso3d_j1_hetero", (9, 12, 20), 28, arrays={"mu": (0.05, 0.1, 0.1), "omega": (0.5, 2.0, 0.0)},
     ell=1.0, dim=3, max_num_iter=20, isotropic=False, solver=J1)
# --- fp32 (the reference computes natively in fp32 on the anisotropic/Jacobi route) ---
case("aniso2d_j1_f32", (37, 52), 29, dtype="float32", mu=0.1, omega=1.0, ell=1.0, dim=2,
     max_num_iter=25, isotropic=False, solver=J1)
case("aniso3d_j1_f32", (13, 17, 11), 30, dtype="float32", mu=0.1, omega=1.0, ell=1.0, dim=3,
     max_num_iter=25, isotropic=False, solver=J1)
# --- convergence / warm start / tol branches ------------------------------------------
case("aniso2d_j1_eps", (37, 52), 31, mu=0.1, omega=1.0, dim=2, max_num_iter=200,
     eps=2e-3, isotropic=False, solver=J1)  # ell = 2*mu: converges after a few iterations
case("iso2d_cg3_eps", (37, 52), 39, mu=0.1, omega=1.0, ell=1.0, dim=2, max_num_iter=200,
     eps=2e-3, isotropic=True, solver={"kind": "cg", "maxiter": 3, "tol": 0.0})
case("aniso2d_jtol", (37, 52), 32, mu=0.1, omega=1.0, ell=1.0, dim=2, max_num_iter=10,
     isotropic=False, solver={"kind": "jacobi", "maxiter": 30, "tol": 1e-3})
case("aniso2d_j1_x0", (24, 31), 33, x0=True, mu=0.1, omega=1.0, ell=1.0, dim=2, max_num_iter=10,
     isotropic=False, solver=J1)
case("iso2d_cg5_x0", (24, 31), 34, x0=True, mu=0.1, omega=1.0, ell=1.0, dim=2, max_num_iter=10,
     isotropic=True, solver={"kind": "cg", "maxiter": 5, "tol": 0.0})
# --- through the public entry point darsia.tvd (defaults: ell = 1.0, eps = 2e-4, Q1/Q17) --
case("tvd_entry_aniso", (37, 52), 35, entry="tvd", method="heterogeneous bregman", weight=0.1,
     omega=1.0, dim=2, isotropic=False, max_num_iter=40, eps=1e-3, solver=J1)
case("tvd_entry_iso_cg", (37, 52), 36, entry="tvd", method="heterogeneous bregman", weight=0.1,
     omega=0.8, dim=2, isotropic=True, max_num_iter=15, eps=None,
     solver={"kind": "cg", "maxiter": 4, "tol": 0.0})

# --- MG V-cycle as the inner solver (utils/linear_solvers/mg.py) ------------------
case("aniso2d_mg", (37, 52), 45, mu=0.1, omega=1.0, ell=1.0, dim=2, max_num_iter=8, isotropic=False,
     solver={"kind": "mg", "depth": 2, "smoother_iterations": 3, "maxiter": 2})
case("iso2d_mg_tol", (40, 33), 46, mu=0.1, omega=0.5, ell=2.0, dim=2, max_num_iter=6, isotropic=True,
     solver={"kind": "mg", "depth": 1, "smoother_iterations": 5, "maxiter": 20, "tol": 1e-3})
case("aniso3d_mg", (13, 18, 11), 47, mu=0.1, omega=1.0, ell=1.0, dim=3, max_num_iter=5, isotropic=False,
     solver={"kind": "mg", "depth": 1, "smoother_iterations": 2, "maxiter": 1})
case("aniso2d_mg_f32", (37, 52), 48, dtype="float32", mu=0.1, omega=1.0, ell=1.0, dim=2, max_num_iter=6,
     isotropic=False, solver={"kind": "mg", "depth": 2, "smoother_iterations": 3, "maxiter": 1})
# MG with per-voxel coefficients (mg.py:134-149, 170-192: restricted on the way down, block averages afterwards),
# a per-voxel mu with scalar omega / ell, and a warm start
case("aniso2d_mg_hetero", (37, 52), 60, arrays={"omega": (0.5, 2.0, 0.0), "ell": (0.5, 1.5, 0.0)}, mu=0.1, dim=2,
     max_num_iter=6, isotropic=False, solver={"kind": "mg", "depth": 2, "smoother_iterations": 3, "maxiter": 2, "hetero": True})
case("iso2d_mg_hetero_omega", (41, 36), 61, arrays={"omega": (0.5, 2.0, 0.0)}, mu=0.1, ell=1.0, dim=2,
     max_num_iter=6, isotropic=True, solver={"kind": "mg", "depth": 1, "smoother_iterations": 4, "maxiter": 1, "hetero": True})
case("aniso3d_mg_hetero", (13, 18, 11), 62, arrays={"omega": (0.5, 2.0, 0.0), "ell": (0.5, 1.5, 0.0)}, mu=0.1, dim=3,
     max_num_iter=4, isotropic=False, solver={"kind": "mg", "depth": 1, "smoother_iterations": 2, "maxiter": 2, "hetero": True})
case("aniso2d_mg_hetero_f32", (37, 52), 63, dtype="float32", arrays={"omega": (0.5, 2.0, 0.0), "ell": (0.5, 1.5, 0.0)}, mu=0.1,
     dim=2, max_num_iter=5, isotropic=False,
     solver={"kind": "mg", "depth": 1, "smoother_iterations": 3, "maxiter": 1, "hetero": True})
case("aniso2d_mg_mu_map", (37, 52), 64, arrays={"mu": (0.05, 0.1, 0.1)}, omega=1.0, ell=1.0, dim=2, max_num_iter=6,
     isotropic=False, solver={"kind": "mg", "depth": 1, "smoother_iterations": 3, "maxiter": 2})
case("aniso2d_mg_x0", (24, 31), 65, x0=True, mu=0.1, omega=1.0, ell=1.0, dim=2, max_num_iter=5, isotropic=False,
     solver={"kind": "mg", "depth": 1, "smoother_iterations": 3, "maxiter": 1})
# --- adaptive ell schedule (split_bregman_tvd.py:149-158) -------------------------------
# parity_tol: ell = 1/|grad u| amplifies round-off wherever the gradient is tiny -- the reference's
# own output moves by 2e-7 (Jacobi) / 1e-5 (CG) / 2e-3 (fp32) rel. L2 when its input is perturbed by
# one ulp (tests/test_oracle_golden.py::test_adaptive_conditioning), so these fixtures carry the
# tolerance their conditioning allows instead of north_star's 1e-10 / 1e-5.
case("aniso2d_j1_adaptive", (37, 52), 40, mu=0.1, omega=1.0, ell=1.0, dim=2, max_num_iter=12,
     isotropic=False, solver=J1, adaptive_every=1, parity_tol=1e-5)
case("aniso2d_j3_adaptive", (30, 41), 41, mu=0.1, omega=1.0, ell=1.0, dim=2, max_num_iter=10,
     isotropic=False, solver={"kind": "jacobi", "maxiter": 3}, adaptive_every=1)
case("aniso2d_cg5_adaptive", (37, 52), 40, mu=0.1, omega=1.0, ell=1.0, dim=2, max_num_iter=8,
     isotropic=False, solver={"kind": "cg", "maxiter": 5, "tol": 0.0}, adaptive_every=3, parity_tol=1e-3)
case("aniso3d_j1_adaptive", (9, 12, 20), 43, mu=0.1, omega=1.0, ell=1.0, dim=3, max_num_iter=10,
     isotropic=False, solver=J1, adaptive_every=1)
case("aniso2d_j1_adaptive_f32", (37, 52), 44, dtype="float32", mu=0.1, omega=1.0, ell=1.0, dim=2,
     max_num_iter=8, isotropic=False, solver=J1, adaptive_every=1, parity_tol=2e-2)
# (the reference's schedule is unstable for most other settings -- e.g. every 2nd iteration with
#  Jacobi, or every iteration with CG, blows up to 1e7..1e100 -- such runs are chaotic, not fixtures)


def run_case(da, c):
    dtype = np.dtype(c["dtype"])
    img = phantom(c["shape"], c["seed"], dtype)
    params = dict(c["params"])
    arrays = {}
    for i, (k, (lo, hi, zf)) in enumerate(sorted(c["arrays"].items())):
        arrays[k] = coeff_field(c["shape"], 1000 + c["seed"] * 10 + i, lo, hi, zf).astype(dtype)
    kwargs = dict(params)
    kwargs.pop("parity_tol", None)
    kwargs["solver"] = make_solver(da, params["solver"], arrays)
    every = kwargs.pop("adaptive_every", None)
    if every:
        kwargs["adaptive"] = lambda it, every=every: it % every == every - 1
    store = {"img": img}
    if c["x0"]:
        rng = np.random.default_rng(5000 + c["seed"])
        dim = params["dim"]
        u0 = np.clip(img + 0.05 * rng.standard_normal(img.shape), 0, 1).astype(dtype)
        d0 = (0.05 * rng.standard_normal((*img.shape, dim))).astype(dtype)
        b0 = (0.05 * rng.standard_normal((*img.shape, dim))).astype(dtype)
        store.update(x0_u=u0, x0_d=d0, x0_b=b0)
        kwargs["x0"] = (u0.copy(), d0.copy(), b0.copy())
    if c["entry"] == "tvd":
        if "mu" in arrays:
            kwargs["weight"] = arrays["mu"]
        if "omega" in arrays:
            kwargs["omega"] = arrays["omega"]
        out = da.tvd(img.copy(), **kwargs)
    else:
        kwargs.update(arrays)
        out = da.split_bregman_tvd(img.copy(), **kwargs)
    store.update({"coef_" + k: v for k, v in arrays.items()})
    store["out"] = out
    store["params"] = np.array(json.dumps({"entry": c["entry"], "x0": c["x0"], **params}))
    return store


def operator_cases(da):
    """Operator-level golden outputs (difference stencils, Laplacian, Jacobi sweep)."""
    store = {}
    for tag, shape, dim in (("2d", (7, 9), 2), ("3d", (5, 4, 6), 3), ("thin", (1, 6), 2)):
        rng = np.random.default_rng({"2d": 101, "3d": 102, "thin": 103}[tag])
        x = rng.standard_normal(shape)
        ell = 0.5 + rng.random(shape)
        store[f"{tag}_x"] = x
        store[f"{tag}_ell"] = ell
        for ax in range(dim):
            store[f"{tag}_bdiff{ax}"] = da.backward_diff(x, ax, dim)
            store[f"{tag}_fdiff{ax}"] = da.forward_diff(x, ax, dim)
        store[f"{tag}_lap_scalar"] = da.laplace(x, dim=dim, diffusion_coeff=0.7)
        store[f"{tag}_lap_array"] = da.laplace(x, dim=dim, diffusion_coeff=ell)
        jac = da.Jacobi(maxiter=1, dim=dim, mass_coeff=1.3, diffusion_coeff=0.7)
        store[f"{tag}_nacc"] = jac._neighbor_accumulation(x)
        rhs = rng.standard_normal(shape)
        store[f"{tag}_rhs"] = rhs
        store[f"{tag}_sweep"] = jac(x0=x, rhs=rhs)
        cg = da.CG(maxiter=4, tol=0.0, dim=dim, mass_coeff=1.3, diffusion_coeff=0.7)
        store[f"{tag}_cg4"] = cg(x0=x, rhs=rhs)
        cgh = da.CG(maxiter=4, tol=0.0, dim=dim, mass_coeff=1.3, diffusion_coeff=ell)
        store[f"{tag}_cg4_hetero"] = cgh(x0=x, rhs=rhs)
        if tag != "thin":
            mg = da.MG(depth=1, smoother_iterations=2, maxiter=2, dim=dim, mass_coeff=1.3, diffusion_coeff=0.7)
            store[f"{tag}_mg_restrict"] = mg.restriction(x)
            store[f"{tag}_mg_prolong"] = mg.prolongation(mg.restriction(x))
            store[f"{tag}_mg_operator_h2"] = mg.operator(x, h=2)
            store[f"{tag}_mg2"] = mg(x0=x.copy(), rhs=rhs)
            om = 0.5 + rng.random(shape)
            store[f"{tag}_om"] = om
            mgh = da.MG(depth=1, smoother_iterations=2, maxiter=2, dim=dim, mass_coeff=om.copy(), diffusion_coeff=ell.copy())
            store[f"{tag}_mg_operator_h2_hetero"] = mgh.operator(x, h=2)
            store[f"{tag}_mg2_hetero"] = mgh(x0=x.copy(), rhs=rhs)
            store[f"{tag}_mg2_hetero_om_after"] = mgh.mass_coeff  # the block averages the solver is left with
    return store


def h1_cases(da):
    """H1 regularisation (restoration/h1_regularization.py) -- SURVEY 8f "next" row 1."""
    store = {}
    vol = phantom((11, 14, 9), 51)
    store["vol"] = vol
    store["vol_jac_tol"] = da.H1_regularization(vol.copy(), mu=1, omega=1, dim=3,
                                                solver=da.Jacobi(maxiter=30, tol=1e-6))  # examples/regularization.py:33-40
    store["vol_cg"] = da.H1_regularization(vol.copy(), mu=0.5, omega=2.0, dim=3, solver=da.CG(maxiter=8, tol=0.0))
    store["vol_mg"] = da.H1_regularization(vol.copy(), mu=1.0, omega=1.0, dim=3,
                                           solver=da.MG(depth=1, smoother_iterations=3, maxiter=4))  # examples/regularization.py:39
    # MG with per-voxel coefficients given to its constructor (mg.py:70-73); H1 passes omega / mu on as mass / diffusion
    vol_om = coeff_field(vol.shape, 53, 0.5, 2.0)
    vol_mu = coeff_field(vol.shape, 54, 0.5, 1.5)
    store["vol_om"], store["vol_mu"] = vol_om, vol_mu
    store["vol_mg_hetero"] = da.H1_regularization(
        vol.copy(), mu=vol_mu.copy(), omega=vol_om.copy(), dim=3,
        solver=da.MG(depth=1, smoother_iterations=3, maxiter=3, dim=3, mass_coeff=vol_om.copy(), diffusion_coeff=vol_mu.copy()))
    rgb = phantom((23, 31, 3), 52)
    store["rgb"] = rgb
    store["rgb_jac5"] = da.H1_regularization(rgb.copy(), mu=0.7, omega=1.0, dim=2, solver=da.Jacobi(maxiter=5))
    store["rgb_jac_tol"] = da.H1_regularization(rgb.copy(), mu=0.7, omega=1.0, dim=2,
                                                solver=da.Jacobi(maxiter=40, tol=1e-3))
    return store


def main():
    da = load_reference()
    only = set(sys.argv[1:])  # optional: regenerate only the named cases
    for c in CASES:
        if only and c["name"] not in only:
            continue
        store = run_case(da, c)
        np.savez_compressed(os.path.join(OUT, c["name"] + ".npz"), **store)
        print(f"{c['name']:28s} shape={c['shape']} out_norm={np.linalg.norm(store['out']):.12g}")
    if not only or "operators" in only:
        np.savez_compressed(os.path.join(OUT, "operators.npz"), **operator_cases(da))
        print("operators.npz written")
    if only and "h1" not in only:
        return
    np.savez_compressed(os.path.join(OUT, "h1.npz"), **h1_cases(da))
    print("h1.npz written")


if __name__ == "__main__":
    main()
