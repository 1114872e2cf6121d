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


def cg_fixed_count_cases():
    """(shape, dim, tile, cg iterations, isotropic, hetero (mu, omega AND ell maps), dtype, batch): CG with
    tol == 0 runs the single-pass kernels (K7); thin / tiny extents put every plane of the two-deep
    pipeline on a boundary, tiny tiles put every halo row and column on a seam."""
    f8, f4 = np.float64, np.float32
    return [
        ((37, 150), 2, (32, 0, 5), 4, True, False, f8, ()),
        ((37, 150), 2, (32, 0, 5), 1, False, False, f8, ()),
        ((37, 150), 2, (32, 0, 5), 2, False, True, f8, (2,)),
        ((11, 19, 70), 3, (8, 6, 3), 5, True, False, f8, ()),
        ((11, 19, 70), 3, (8, 7, 4), 3, False, True, f4, ()),
        ((9, 5, 6), 3, (0, 0, 0), 6, False, False, f8, ()),
        ((3, 40), 2, (0, 0, 0), 6, False, False, f8, ()),
        ((1, 40), 2, (0, 0, 0), 3, False, False, f8, ()),
        ((2, 1, 7), 3, (0, 0, 0), 3, False, False, f8, ()),
    ]


def run_cg_fixed_pair(da, oracle, shape, dim, tile, m, iso, het, dtype, batch, fused=True):
    img = phantom(tuple(shape) + tuple(batch), 5, dtype)
    kw = dict(mu=0.1, omega=1.0, ell=1.0)
    if het:
        kw["mu"] = coeff_field(shape, 8, 0.05, 0.1, 0.1).astype(dtype)
        kw["omega"] = coeff_field(shape, 9, 0.5, 2.0).astype(dtype)
        kw["ell"] = coeff_field(shape, 10, 0.5, 1.5).astype(dtype)
    ref = oracle.split_bregman_tvd(img, dim=dim, max_num_iter=6, isotropic=iso, solver=oracle.CG(maxiter=m, tol=0.0), **kw)
    out = da.split_bregman_tvd(img, dim=dim, max_num_iter=6, isotropic=iso, solver=da.CG(maxiter=m, tol=0.0),
                               tile=tile, fused=(2 if fused else False), **kw)  # 2: single-pass also in 3-D
    return out, ref


def jacobi_blocked_cases():
    """(shape, dim, tile, sweeps, isotropic, dtype, batch): Jacobi(maxiter=K >= 2) with scalar coefficients
    runs temporally blocked passes (4 / 2 sweeps per pass); thin extents and tiny tiles put every stage of the
    pipeline on a boundary or a seam."""
    f8, f4 = np.float64, np.float32
    return [
        ((37, 150), 2, (32, 0, 5), 20, False, f8, ()),
        ((37, 150), 2, (32, 0, 7), 6, True, f8, (2,)),
        ((37, 150), 2, (32, 0, 5), 3, False, f8, ()),
        ((37, 150), 2, (32, 0, 5), 5, False, f4, ()),
        ((11, 19, 70), 3, (8, 8, 3), 5, False, f8, ()),
        ((11, 19, 70), 3, (8, 10, 4), 4, True, f4, ()),
        ((9, 5, 6), 3, (0, 0, 0), 7, False, f8, ()),
        ((3, 40), 2, (0, 0, 0), 9, False, f8, ()),
        ((1, 40), 2, (0, 0, 0), 6, False, f8, ()),
        ((2, 1, 7), 3, (0, 0, 0), 3, False, f8, ()),
        ((70, 300), 2, (0, 0, 0), 20, False, f8, ()),
    ]


def run_jacobi_blocked_pair(da, oracle, shape, dim, tile, K, iso, dtype, batch):
    img = phantom(tuple(shape) + tuple(batch), 5, dtype)
    kw = dict(mu=0.1, omega=0.5, ell=1.0, dim=dim, max_num_iter=4, isotropic=iso)
    ref = oracle.split_bregman_tvd(img, solver=oracle.Jacobi(maxiter=K), **kw)
    out = da.split_bregman_tvd(img, solver=da.Jacobi(maxiter=K), tile=tile, fused=2, **kw)  # 2: blocked in 3-D too
    one = da.split_bregman_tvd(img, solver=da.Jacobi(maxiter=K), tile=tile, fused=False, **kw)
    return out, one, ref


def jacobi1_pair_cases():
    """(shape, tile, outer iterations, isotropic, hetero, dtype, batch): Jacobi(maxiter=1) without eps in 2-D
    runs TWO outer iterations per launch (K1x2, a four-stage pipeline); odd counts end with a single one."""
    f8, f4 = np.float64, np.float32
    return [
        ((37, 150), (32, 0, 5), 6, False, False, f8, ()),
        ((37, 150), (32, 0, 7), 7, True, False, f8, (2,)),
        ((37, 150), (32, 0, 5), 5, False, True, f8, ()),
        ((37, 150), (32, 0, 5), 6, True, True, f4, ()),
        ((3, 40), (0, 0, 0), 6, False, False, f8, ()),
        ((1, 40), (0, 0, 0), 4, False, False, f8, ()),
        ((2, 7), (0, 0, 0), 5, True, False, f8, ()),
        ((70, 300), (0, 0, 0), 9, False, False, f8, ()),
        ((5, 9), (0, 0, 0), 2, False, False, f8, ()),
    ]


def run_jacobi1_pair(da, oracle, shape, tile, nit, iso, het, dtype, batch):
    img = phantom(tuple(shape) + tuple(batch), 5, dtype)
    kw = dict(mu=0.1, omega=0.5, ell=1.0)
    if het:
        kw["mu"] = coeff_field(shape, 8, 0.05, 0.1, 0.1).astype(dtype)
        kw["omega"] = coeff_field(shape, 9, 0.5, 2.0).astype(dtype)
        kw["ell"] = coeff_field(shape, 10, 0.5, 1.5).astype(dtype)
    hist = []
    ref = oracle.split_bregman_tvd(img, dim=2, max_num_iter=nit, isotropic=iso, solver=oracle.Jacobi(maxiter=1),
                                   history=hist, **kw)
    out, info = da.split_bregman_tvd(img, dim=2, max_num_iter=nit, isotropic=iso, solver=da.Jacobi(maxiter=1), tile=tile,
                                     return_info=True, fused=2, **kw)  # 2: pairs for every variant, not only where they win
    one = da.split_bregman_tvd(img, dim=2, max_num_iter=nit, isotropic=iso, solver=da.Jacobi(maxiter=1), tile=tile,
                               fused=3, **kw)  # 3: one iteration per launch
    return out, one, ref, info, hist


def chambolle_cases():
    """(shape, dtype, weight, eps, max_num_iter) for ``method="chambolle"`` (every axis spatial)."""
    f8, f4 = np.float64, np.float32
    return [
        ((37, 52), f8, 0.1, 2e-4, 200),
        ((37, 150), f8, 0.3, 1e-5, 60),
        ((11, 19, 33), f8, 0.1, 2e-4, 200),
        ((37, 52), f4, 0.1, 2e-4, 200),
        ((3, 40), f8, 0.1, 2e-4, 30),
        ((1, 40), f8, 0.1, 2e-4, 10),
        ((2, 1, 7), f8, 0.1, 0.0, 5),
        ((24, 40, 3), f8, 0.2, 2e-4, 1),
    ]


def check_mg_standalone(da, tag, dim):
    """darsia_b200.MG standalone (as H1_regularization would use it) against the reference's MG outputs in
    operators.npz: scalar coefficients, the refused cases, and per-voxel coefficients given to the constructor
    (mg.py:70-73: restricted on the way down, block averages afterwards -- which the solver object is left holding,
    mg.py:134-149, 170-192)."""
    import os

    import pytest

    golden = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "operators.npz")
    rel = lambda a, b: float(np.linalg.norm(a - b) / np.linalg.norm(b))  # noqa: E731
    with np.load(golden) as z:
        x, rhs, ref = z[f"{tag}_x"], z[f"{tag}_rhs"], z[f"{tag}_mg2"]
        om, ell, refh, om_after = z[f"{tag}_om"], z[f"{tag}_ell"], z[f"{tag}_mg2_hetero"], z[f"{tag}_mg2_hetero_om_after"]
    mg = da.MG(depth=1, smoother_iterations=2, maxiter=2, dim=dim, mass_coeff=1.3, diffusion_coeff=0.7)
    assert rel(mg(x.copy(), rhs), ref) <= 1e-13
    with pytest.raises(ValueError):  # a coarse grid would be empty (the reference crashes there)
        da.MG(depth=3, dim=dim, mass_coeff=1.0, diffusion_coeff=1.0)(x, rhs)
    mgh = da.MG(depth=1, smoother_iterations=2, maxiter=2, dim=dim, mass_coeff=om.copy(), diffusion_coeff=ell.copy())
    assert rel(mgh(x.copy(), rhs), refh) <= 1e-13
    np.testing.assert_allclose(mgh.mass_coeff, om_after, rtol=0, atol=1e-15)
    late = da.MG(depth=1, dim=dim)  # heterogeneous was decided at construction: the reference fails on the coarse grid
    late.update_params(mass_coeff=om, diffusion_coeff=1.0)
    with pytest.raises(ValueError):
        late(x, rhs)


def check_mg_warm_start(da, oracle):
    """MG as the inner solver: 7 iterations == 4 iterations, (u, d, b) handed over as x0, 3 more; then batched with a
    tolerance (one problem at a time) from an arbitrary (not self-consistent) warm start, against the oracle."""
    rel = lambda a, b: float(np.linalg.norm(a - b) / np.linalg.norm(b))  # noqa: E731
    img = phantom((24, 31, 2), 3)
    kw = dict(mu=0.1, omega=0.8, ell=1.0, dim=2, isotropic=True)
    mk = lambda m, **k: m.MG(depth=1, smoother_iterations=2, maxiter=2, **k)  # noqa: E731
    ref7 = oracle.split_bregman_tvd(img, max_num_iter=7, solver=mk(oracle), **kw)
    u4, d4, b4 = da.split_bregman_tvd(img, max_num_iter=4, solver=mk(da), return_state=True, **kw)
    assert d4.shape == (*img.shape, 2) and b4.shape == (*img.shape, 2)
    out = da.split_bregman_tvd(img, max_num_iter=3, solver=mk(da), x0=(u4, d4, b4), **kw)
    assert rel(out, ref7) <= 1e-13
    rng = np.random.default_rng(1)
    x0 = (np.clip(img + 0.05 * rng.standard_normal(img.shape), 0, 1), 0.05 * rng.standard_normal((*img.shape, 2)),
          0.05 * rng.standard_normal((*img.shape, 2)))
    ref = oracle.split_bregman_tvd(img, max_num_iter=3, solver=mk(oracle, tol=1e-3), x0=tuple(v.copy() for v in x0), **kw)
    out, d, b = da.split_bregman_tvd(img, max_num_iter=3, solver=mk(da, tol=1e-3), x0=x0, return_state=True, **kw)
    assert rel(out, ref) <= 1e-13 and d.shape == b.shape == (*img.shape, 2)
    # verbose after a warm start: "starts at" is _functional(img), the first iteration's line carries the warm start's energy
    import contextlib
    import io
    import re

    one, x0_one = img[..., 0].copy(), (x0[0][..., 0].copy(), x0[1][..., 0, :].copy(), x0[2][..., 0, :].copy())
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        da.split_bregman_tvd(one, max_num_iter=2, solver=mk(da), verbose=True, x0=x0_one, **kw)
    got = [float(v) for v in re.findall(r"[-+]?\d+\.\d+(?:[eE][-+]?\d+)?", buf.getvalue())]
    assert len(got) == 5  # start energy + 2 x (increment, energy)
    # split_bregman_tvd.py:103-104 prints _functional(img); the line of iteration 0 carries the energy of the warm start
    np.testing.assert_allclose(got[0], oracle.energy_functional(one, one, kw["mu"], kw["omega"], 2), rtol=1e-12)
    np.testing.assert_allclose(got[2], oracle.energy_functional(x0_one[0], one, kw["mu"], kw["omega"], 2), rtol=1e-12)


def mg_hetero_cases():
    """(shape, dim, depth, isotropic, dtype, batch, which maps): MG with per-voxel coefficients across several tiles and
    odd extents on every level (the restriction drops a row / column, the prolongation pads it back)."""
    f8, f4 = np.float64, np.float32
    return [
        ((75, 131), 2, 2, False, f8, (), ("omega", "ell")),
        ((75, 131), 2, 1, True, f8, (2,), ("omega",)),
        ((21, 30, 27), 3, 1, False, f8, (), ("omega", "ell")),
        ((64, 96), 2, 2, False, f4, (), ("ell",)),
    ]


def run_mg_hetero_pair(da, oracle, shape, dim, depth, iso, dtype, batch, which):
    img = phantom(tuple(shape) + tuple(batch), 5, dtype)
    kw = dict(mu=0.1, omega=0.8, ell=1.2)
    if "omega" in which:
        kw["omega"] = coeff_field(shape, 9, 0.5, 2.0).astype(dtype)
    if "ell" in which:
        kw["ell"] = coeff_field(shape, 10, 0.5, 1.5).astype(dtype)
    mk = lambda m: m.MG(depth=depth, smoother_iterations=3, maxiter=2, dim=dim,  # noqa: E731
                        mass_coeff=(kw["omega"].copy() if "omega" in which else None),
                        diffusion_coeff=(kw["ell"].copy() if "ell" in which else None))
    ref = oracle.split_bregman_tvd(img, dim=dim, max_num_iter=4, isotropic=iso, solver=mk(oracle), **kw)
    out = da.split_bregman_tvd(img, dim=dim, max_num_iter=4, isotropic=iso, solver=mk(da), **kw)
    return out, ref
