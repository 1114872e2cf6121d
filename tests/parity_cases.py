"""Seeded parity cases shared by the emulator (CPU) and the GPU test modules."""

import itertools

import numpy as np

from golden.generate_golden import coeff_field, phantom

SOLVERS = {
    "j1": lambda api: api.Jacobi(maxiter=1),
    "j3": lambda api: api.Jacobi(maxiter=3),
    "j20": lambda api: api.Jacobi(maxiter=20),
    "jtol": lambda api: api.Jacobi(maxiter=10, tol=1e-2),
    "cg4": lambda api: api.CG(maxiter=4, tol=0.0),
    "cg10": lambda api: api.CG(maxiter=10, tol=0.0),
    "cgtol": lambda api: api.CG(maxiter=40, tol=1e-3),
}


def make_inputs(shape, batch=(), seed=7, dtype=np.float64, hetero=False):
    img = phantom(tuple(shape) + tuple(batch), seed, dtype)
    kw = dict(mu=0.1, omega=1.0, ell=1.0)
    if hetero:
        kw["mu"] = coeff_field(shape, seed + 1, 0.05, 0.1, 0.1).astype(dtype)
        kw["omega"] = coeff_field(shape, seed + 2, 0.5, 2.0).astype(dtype)
    return img, kw


def run_pair(da, oracle, shape, dim, iso, solver, batch=(), dtype=np.float64, hetero=False, nit=6, **extra):
    """Same call through the product API and the oracle; returns (out, ref)."""
    img, kw = make_inputs(shape, batch, dtype=dtype, hetero=hetero)
    ref = oracle.split_bregman_tvd(img, dim=dim, max_num_iter=nit, isotropic=iso, solver=SOLVERS[solver](oracle), **kw)
    out = da.split_bregman_tvd(img, dim=dim, max_num_iter=nit, isotropic=iso, solver=SOLVERS[solver](da), **kw, **extra)
    return out, ref


def small_grid():
    """(shape, dim, tile) x options: small shapes with forced tiny tiles -> every halo/chunk seam."""
    geoms = [((37, 150), 2, (32, 0, 5)), ((11, 19, 70), 3, (8, 6, 3))]
    opts = itertools.product((False, True), ("j1", "j3", "cg4", "jtol"), (False, True), (np.float64, np.float32))
    for (shape, dim, tile), (iso, solver, hetero, dtype) in itertools.product(geoms, opts):
        yield shape, dim, tile, iso, solver, hetero, dtype


def adaptive_cases():
    """(shape, dim, tile, solver, isotropic): the adaptive ell schedule (every iteration) across tile and
    chunk seams, 3-D, isotropic and batched; all well-conditioned (checked against the oracle at 1e-12)."""
    return [
        ((11, 19, 70), 3, (8, 6, 3), "j3", False),
        ((37, 150), 2, (32, 0, 5), "j3", False),
        ((11, 19, 70), 3, (8, 6, 3), "j1", True),
        ((21, 70, 2), 2, (32, 0, 5), "j3", False),
    ]


def run_adaptive_pair(da, oracle, shape, dim, tile, solver, iso):
    img = phantom(tuple(shape), 5)
    kw = dict(mu=0.1, omega=1.0, ell=1.0, dim=dim, max_num_iter=6, isotropic=iso, adaptive=lambda it: True)
    ref = oracle.split_bregman_tvd(img, solver=SOLVERS[solver](oracle), **kw)
    out = da.split_bregman_tvd(img, solver=SOLVERS[solver](da), tile=tile, **kw)
    return out, ref
