"""CPU: the kernel SOURCE (darsia_b200/csrc/tvdk_core.h) stepped thread by thread on the host
(tests/emu) against the golden vectors of the reference and against the oracle.

This checks the index / halo / boundary logic of every kernel, the host driver and the Python
boundary without a GPU.  It says nothing about the sm_100a build -- the ``-m gpu`` modules do.
"""

import numpy as np
import pytest

import fuzz_cases
import oracle
from conftest import golden_case_names, load_golden, rel_l2, run_case_with, tolerance_for
from emu.harness import emulated
from parity_cases import (adaptive_cases, cg_fixed_count_cases, check_mg_standalone, check_mg_warm_start, mg_hetero_cases, run_mg_hetero_pair, chambolle_cases, jacobi1_pair_cases, jacobi_blocked_cases,
                          run_adaptive_pair, run_cg_fixed_pair, run_jacobi1_pair, run_jacobi_blocked_pair, run_pair,
                          small_grid)

from tma_cases import run_stencil_case, run_tma2d_case, run_tma_case, stencil_cases, tma2d_cases, tma_cases

import darsia_b200 as da


@pytest.fixture(scope="module", autouse=True)
def _emu():
    with emulated() as lib:
        assert lib.is_emulator
        yield lib


@pytest.mark.parametrize("name", golden_case_names())
def test_golden(name):
    arrays, params = load_golden(name)
    out = run_case_with(da, arrays, params)
    ref = arrays["out"]
    assert out.shape == ref.shape and out.dtype == ref.dtype
    assert rel_l2(out, ref) <= tolerance_for(ref.dtype, params)


@pytest.mark.parametrize("shape,dim,tile,iso,solver,hetero,dtype", list(small_grid()))
def test_tile_seams(shape, dim, tile, iso, solver, hetero, dtype):
    out, ref = run_pair(da, oracle, shape, dim, iso, solver, dtype=dtype, hetero=hetero, tile=tile)
    assert rel_l2(out, ref) <= (1e-12 if dtype == np.float64 else 2e-5)


@pytest.mark.parametrize("shape,dim,tile,m,iso,het,dtype,batch", cg_fixed_count_cases())
@pytest.mark.parametrize("fused", [True, False])
def test_cg_fixed_count_single_pass_and_two_pass(shape, dim, tile, m, iso, het, dtype, batch, fused):
    """CG(tol=0): the single-pass route (fused=True, K7) and the two-kernel route agree with the oracle."""
    out, ref = run_cg_fixed_pair(da, oracle, shape, dim, tile, m, iso, het, dtype, batch, fused=fused)
    assert rel_l2(out, ref) <= (1e-13 if dtype == np.float64 else 2e-6)


@pytest.mark.parametrize("shape,dim,tile,K,iso,dtype,batch", jacobi_blocked_cases())
def test_jacobi_sweeps_temporally_blocked(shape, dim, tile, K, iso, dtype, batch):
    """Jacobi(K): S sweeps per pass (K4xS) == one sweep per pass, bitwise (same arithmetic per voxel), and
    both agree with the oracle."""
    out, one, ref = run_jacobi_blocked_pair(da, oracle, shape, dim, tile, K, iso, dtype, batch)
    np.testing.assert_array_equal(out, one)
    assert rel_l2(out, ref) <= (1e-13 if dtype == np.float64 else 2e-6)


@pytest.mark.parametrize("shape,tile,nit,iso,het,dtype,batch", jacobi1_pair_cases())
def test_jacobi1_two_iterations_per_launch(shape, tile, nit, iso, het, dtype, batch):
    """K1x2 == K1 applied twice, bitwise; both agree with the oracle, and the per-iteration increments of
    the paired launches are the oracle's."""
    out, one, ref, info, hist = run_jacobi1_pair(da, oracle, shape, tile, nit, iso, het, dtype, batch)
    np.testing.assert_array_equal(out, one)
    assert rel_l2(out, ref) <= (1e-13 if dtype == np.float64 else 2e-6)
    assert list(info.iters) == [nit] * max(1, int(np.prod(batch, dtype=int)))
    for b, h in enumerate(hist if batch else [hist]):
        np.testing.assert_allclose(info.history[: len(h), b], h, rtol=(1e-9 if dtype == np.float64 else 1e-3))


@pytest.mark.parametrize("shape,dim,tile,solver,iso", adaptive_cases())
def test_adaptive_schedule(shape, dim, tile, solver, iso):
    out, ref = run_adaptive_pair(da, oracle, shape, dim, tile, solver, iso)
    assert rel_l2(out, ref) <= 1e-12


@pytest.mark.parametrize("fused", [True, False])
def test_batched_channels_are_independent(fused):
    out, ref = run_pair(da, oracle, (20, 70), 2, True, "j1", batch=(3,), tile=(32, 0, 6), fused=fused)
    assert rel_l2(out, ref) <= 1e-12
    out, ref = run_pair(da, oracle, (7, 12, 20), 3, False, "cg4", batch=(2,), tile=(8, 6, 3))
    assert rel_l2(out, ref) <= 1e-12


def test_eps_stops_each_problem_at_the_reference_iteration():
    from golden.generate_golden import phantom

    img = np.stack([phantom((24, 40), 3), 0.3 * phantom((24, 40), 4) + 0.3], axis=-1)
    hist = []
    ref = oracle.split_bregman_tvd(img, mu=0.1, ell=1.0, max_num_iter=60, eps=5e-3, history=hist,
                                   solver=oracle.CG(maxiter=3, tol=0.0))
    out, info = da.split_bregman_tvd(img, mu=0.1, ell=1.0, max_num_iter=60, eps=5e-3, return_info=True,
                                     solver=da.CG(maxiter=3, tol=0.0))
    assert [len(h) for h in hist] == list(info.iters)
    assert len(set(info.iters)) == 2  # the two channels really stop at different iterations
    assert rel_l2(out, ref) <= 1e-12
    for b, h in enumerate(hist):
        np.testing.assert_allclose(info.history[: len(h), b], h, rtol=1e-9)


@pytest.mark.parametrize("name", ["vol_jac_tol", "vol_cg", "rgb_jac5", "rgb_jac_tol", "vol_mg", "vol_mg_hetero"])
def test_h1_regularization_golden(name):
    import h1_cases

    out, ref = h1_cases.run(da, h1_cases.load(), name)
    assert out.shape == ref.shape and rel_l2(out, ref) <= 1e-12


def test_state_round_trip_and_device_resident_io():
    """``return_state`` hands back the reference's ``x0`` tuple (u, d, b): 4 iterations + a warm start of 3
    more == 7 iterations straight.  Tensors in -> tensors out (no host copies), same numbers."""
    import torch
    from golden.generate_golden import coeff_field, phantom

    img = phantom((24, 31, 2), 3)
    mu = coeff_field((24, 31), 8, 0.05, 0.1, 0.1)
    kw = dict(omega=1.0, ell=1.0, dim=2, isotropic=True)
    ref7 = oracle.split_bregman_tvd(img, mu=mu, max_num_iter=7, solver=oracle.CG(maxiter=4, tol=0.0), **kw)
    u4, d4, b4 = da.split_bregman_tvd(img, mu=mu, max_num_iter=4, solver=da.CG(maxiter=4, tol=0.0), return_state=True, **kw)
    assert d4.shape == (*img.shape, 2) and b4.shape == (*img.shape, 2)
    out = da.split_bregman_tvd(img, mu=mu, max_num_iter=3, solver=da.CG(maxiter=4, tol=0.0), x0=(u4, d4, b4), **kw)
    assert rel_l2(out, ref7) <= 1e-13
    ti, tmu = torch.from_numpy(img), torch.from_numpy(mu)
    st = da.split_bregman_tvd(ti, mu=tmu, max_num_iter=4, solver=da.CG(maxiter=4, tol=0.0), return_state=True, **kw)
    assert all(isinstance(a, torch.Tensor) for a in st)
    tout = da.split_bregman_tvd(ti, mu=tmu, max_num_iter=3, solver=da.CG(maxiter=4, tol=0.0), x0=st, **kw)
    assert isinstance(tout, torch.Tensor) and np.array_equal(tout.numpy(), out)
    tv = da.tvd(ti, method="heterogeneous bregman", weight=0.1, max_num_iter=3, eps=None, dim=2)
    assert isinstance(tv, torch.Tensor) and tv.shape == ti.shape

from conftest import GOLDEN_DIR  # noqa: E402


@pytest.mark.parametrize("tag,dim", [("2d", 2), ("3d", 3)])
def test_mg_standalone_and_errors(tag, dim):
    check_mg_standalone(da, tag, dim)


def test_mg_warm_start_and_state_round_trip():
    check_mg_warm_start(da, oracle)


@pytest.mark.parametrize("shape,dim,depth,iso,dtype,batch,which", mg_hetero_cases())
def test_mg_per_voxel_coefficients(shape, dim, depth, iso, dtype, batch, which):
    out, ref = run_mg_hetero_pair(da, oracle, shape, dim, depth, iso, dtype, batch, which)
    assert rel_l2(out, ref) <= (1e-12 if dtype == np.float64 else 2e-5)


def test_mg_inner_solver_batched_and_stopping():
    """MG inside split_bregman_tvd: channels are independent problems (batched without tol/eps, one at a time
    with them) and the outer eps / MG tol stop each problem where the oracle stops it."""
    from golden.generate_golden import phantom

    img = phantom((24, 40, 2), 3)
    kw = dict(mu=0.1, omega=1.0, ell=1.0, dim=2, max_num_iter=5)
    mk = lambda m, **k: m.MG(depth=1, smoother_iterations=2, maxiter=3, **k)  # noqa: E731
    ref = oracle.split_bregman_tvd(img, solver=mk(oracle), **kw)
    assert rel_l2(da.split_bregman_tvd(img, solver=mk(da), **kw), ref) <= 1e-13
    hist = []
    kw2 = dict(kw, max_num_iter=40, eps=1e-2)
    ref = oracle.split_bregman_tvd(img, solver=mk(oracle, tol=1e-2), history=hist, **kw2)
    out, info = da.split_bregman_tvd(img, solver=mk(da, tol=1e-2), return_info=True, **kw2)
    assert [len(h) for h in hist] == list(info.iters)
    assert rel_l2(out, ref) <= 1e-13


def test_series_sub_batches_overlapped():
    """tvd_series: sub-batches solved from two host threads (one stream each on the GPU) == one after the other
    == the oracle per frame."""
    from golden.generate_golden import phantom

    series = np.stack([phantom((24, 40), 10 + t) for t in range(5)], axis=-1)
    kw = dict(dim=2, mu=0.1, omega=1.0, ell=1.0, max_num_iter=6, isotropic=True, batch=2)
    a, rng_a = da.tvd_series(series, overlap=True, **kw)
    b, rng_b = da.tvd_series(series, overlap=False, **kw)
    assert rng_a == rng_b == (0, 5)
    np.testing.assert_array_equal(a, b)
    for t in range(5):
        ref = oracle.split_bregman_tvd(np.ascontiguousarray(series[..., t]), dim=2, mu=0.1, omega=1.0, ell=1.0,
                                       max_num_iter=6, isotropic=True)
        assert rel_l2(a[..., t], ref) <= 1e-13


@pytest.mark.parametrize("shape,dtype,weight,eps,nit", chambolle_cases())
def test_chambolle_matches_its_restatement(shape, dtype, weight, eps, nit):
    """method="chambolle" (scikit-image's algorithm; PARITY UNPINNED -- the checker is a restatement of the
    published algorithm, not scikit-image itself): same result, same stopping iteration, same energies."""
    from golden.generate_golden import phantom

    from darsia_b200.restoration.chambolle import denoise_tv_chambolle

    img = phantom(shape, 5, dtype)
    ref, ri = oracle.denoise_tv_chambolle(img, weight, eps, nit, return_info=True)
    out, oi = denoise_tv_chambolle(img, weight, eps, nit, return_info=True)
    assert out.dtype == img.dtype and out.shape == img.shape
    assert oi["iterations"] == ri["iterations"]
    assert rel_l2(out, ref) <= (1e-13 if dtype == np.float64 else 1e-6)
    np.testing.assert_allclose(oi["energies"], ri["energies"], rtol=(1e-12 if dtype == np.float64 else 1e-5))


def test_tvd_default_method_is_chambolle():
    """darsia.tvd(img) defaults to method="chambolle", weight=1.0 (tvd.py:132-139); properties that hold for the
    algorithm independently of any implementation: one iteration returns the image, the mean is preserved
    (out = image + div p and div p sums to zero), the total variation drops."""
    from golden.generate_golden import phantom

    img = phantom((40, 56), 12)
    out = da.tvd(img, weight=0.2)
    assert rel_l2(out, oracle.denoise_tv_chambolle(img, 0.2, 2e-4, 200)) <= 1e-13
    np.testing.assert_array_equal(da.tvd(img, weight=0.2, max_num_iter=1), img)
    assert abs(out.mean() - img.mean()) <= 1e-14
    tv = lambda x: np.abs(np.diff(x, axis=0)).sum() + np.abs(np.diff(x, axis=1)).sum()  # noqa: E731
    assert tv(out) < 0.6 * tv(img)
    with pytest.raises(NotImplementedError):
        da.tvd(img, method="isotropic bregman")


@pytest.mark.parametrize("case", fuzz_cases.draw(seed=11, n=40), ids=lambda c: f"{c['kind']}-{c['dim']}d-f{c['fused']}")
def test_random_problems_over_every_route(case):
    """Seeded random shapes / tiles / batches / dtypes / coefficient maps / solvers / kernel routes vs the oracle."""
    out, ref = fuzz_cases.run(da, oracle, case)
    assert rel_l2(out, ref) <= (1e-12 if case["dtype"] == np.float64 else 1e-5)  # north_star's bars are 1e-10 / 1e-5


def test_single_pass_cg_predicted_rho_diagnostic():
    """K7 predicts rho = r.r of the next CG iteration from the sums of the current one; the largest relative
    deviation from the directly summed value is recorded per problem (tvdk_stat_t.cg_pred_dev)."""
    from golden.generate_golden import phantom

    img = phantom((60, 200, 2), 4)
    out, info = da.split_bregman_tvd(img, mu=0.1, omega=1.0, ell=1.0, isotropic=True, max_num_iter=20,
                                     solver=da.CG(maxiter=10, tol=0.0), return_info=True)
    assert info.cg_pred_dev.shape == (2,) and np.all(info.cg_pred_dev > 0) and np.all(info.cg_pred_dev < 1e-12)
    ref = oracle.split_bregman_tvd(img, mu=0.1, omega=1.0, ell=1.0, isotropic=True, max_num_iter=20,
                                   solver=oracle.CG(maxiter=10, tol=0.0))
    assert rel_l2(out, ref) <= 1e-13


@pytest.mark.parametrize("iso", [False, True])
@pytest.mark.parametrize("shape,batch,tile,dtype,nit", tma_cases())
def test_tma_staged_fused_jacobi1_3d(shape, batch, tile, dtype, nit, iso):
    """KFusedJ1T (3-D anisotropic Jacobi(1), scalar coefficients): the TMA boxes are emulated by a zero-filling copy."""
    out, ref, info, hist = run_tma_case(da, oracle, shape, batch, tile, dtype, nit, iso)
    assert rel_l2(out, ref) <= (1e-13 if dtype == np.float64 else 2e-6)
    nb = max(1, int(np.prod(batch, dtype=int)))
    assert list(info.iters) == [nit] * nb
    np.testing.assert_allclose(info.history[:nit].reshape(-1), np.array(hist).reshape(nb, nit).T.reshape(-1),
                               rtol=1e-9 if dtype == np.float64 else 1e-3)


@pytest.mark.parametrize("shape,batch,tile,dtype,nit", tma2d_cases())
def test_tma_staged_fused_jacobi1_2d(shape, batch, tile, dtype, nit):
    """KFusedJ1T<T, false, 2> (2-D anisotropic Jacobi(1), scalar coefficients; fused = 4) against the oracle and the
    march kernel."""
    out, ref, info, hist, march = run_tma2d_case(da, oracle, shape, batch, tile, dtype, nit)
    assert rel_l2(out, ref) <= (1e-13 if dtype == np.float64 else 2e-6)
    assert rel_l2(out, march) <= (1e-14 if dtype == np.float64 else 1e-6)
    nb = max(1, int(np.prod(batch, dtype=int)))
    assert list(info.iters) == [nit] * nb
    np.testing.assert_allclose(info.history[:nit].reshape(-1), np.array(hist).reshape(nb, nit).T.reshape(-1),
                               rtol=1e-9 if dtype == np.float64 else 1e-3)


def test_series_streaming_pipeline():
    """tvd_series_stream: frames from a callable / an iterator through the staging rings == frame-by-frame oracle; partial
    last sub-batch, colour channels, shared per-pixel maps, sink delivery in order."""
    from golden.generate_golden import coeff_field, phantom

    shape = (18, 25)
    frames = [phantom(shape, 40 + k) for k in range(7)]
    mu = 0.1 * coeff_field(shape, 5, 0.5, 1.0, 0.1)
    om = coeff_field(shape, 6, 0.5, 2.0)
    got, rng = da.tvd_series_stream(lambda k: frames[k], n_frames=7, dim=2, batch=3, mu=mu, omega=om, ell=1.0, max_num_iter=6,
                                    solver=da.Jacobi(maxiter=1))
    assert rng == (0, 7) and len(got) == 7
    for k in range(7):
        ref = oracle.split_bregman_tvd(frames[k], mu=mu, omega=om, ell=1.0, dim=2, max_num_iter=6, solver=oracle.Jacobi(maxiter=1))
        assert rel_l2(got[k], ref) <= 1e-13
    # this rank's block of a sharded series
    got1, rng1 = da.tvd_series_stream(lambda k: frames[k], n_frames=7, dim=2, batch=2, rank=1, world=2, mu=mu, omega=om, ell=1.0,
                                      max_num_iter=6, solver=da.Jacobi(maxiter=1))
    assert rng1 == da.frame_range(7, 1, 2) and all(np.array_equal(a, b) for a, b in zip(got1, got[rng1[0]:rng1[1]]))
    rgb = [phantom(shape + (3,), 50 + k) for k in range(4)]
    seen = []
    da.tvd_series_stream(iter(rgb), dim=2, batch=2, sink=lambda k, f: seen.append((k, f.copy())), mu=0.1, ell=1.0, isotropic=True,
                         max_num_iter=4, solver=da.CG(maxiter=3, tol=0.0))
    assert [k for k, _ in seen] == [0, 1, 2, 3]
    for k, f in seen:
        ref = oracle.split_bregman_tvd(rgb[k], mu=0.1, ell=1.0, dim=2, isotropic=True, max_num_iter=4, solver=oracle.CG(maxiter=3, tol=0.0))
        assert rel_l2(f, ref) <= 1e-13
    with pytest.raises(AssertionError):
        da.tvd_series_stream(iter([frames[0], np.zeros((5, 5))]), dim=2, mu=0.1, ell=1.0, max_num_iter=1)


@pytest.mark.parametrize("solver", ["cg", "j3"])
@pytest.mark.parametrize("shape,tile,dtype", stencil_cases())
def test_tma_staged_stencil_3d(shape, tile, dtype, solver):
    """KStencilT (3-D cg_pq / Jacobi sweep with scalar coefficients) against the oracle."""
    out, ref = run_stencil_case(da, oracle, shape, tile, dtype, solver)
    assert rel_l2(out, ref) <= (1e-13 if dtype == np.float64 else 2e-6)
