"""GPU (B200): the sm_100a kernels, called through the C ABI, against the golden vectors of
the reference and against the oracle on seeded inputs.  Tolerances are BASELINE.json's:
relative L2 <= 1e-10 in fp64, <= 1e-5 in fp32 -- the tests assert far tighter fp64 bounds
because the arithmetic is the reference's up to reciprocal-multiplies and FMA contraction.
"""

import numpy as np
import pytest

import fuzz_cases
import oracle
from conftest import golden_case_names, load_golden, rel_l2, run_case_with, tolerance_for
from parity_cases import (adaptive_cases, cg_fixed_count_cases, chambolle_cases, jacobi1_pair_cases, jacobi_blocked_cases,
                          run_adaptive_pair, run_cg_fixed_pair, run_jacobi1_pair, run_jacobi_blocked_pair, run_pair,
                          small_grid)

from tma_cases import run_stencil_case, run_tma2d_case, run_tma_case, stencil_cases, tma2d_cases, tma_cases

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def da():
    import torch

    assert torch.cuda.is_available(), "the gpu tests need a CUDA device"
    import darsia_b200
    from darsia_b200 import _cabi

    lib = _cabi.get_library()
    assert not lib.is_emulator and lib.path.endswith("darsia_b200/csrc/libtvdk.so")
    assert lib.device_ok(), lib.last_error()
    return darsia_b200


@pytest.mark.parametrize("name", golden_case_names())
def test_golden(da, name):
    arrays, params = load_golden(name)
    out = run_case_with(da, arrays, params)
    ref = arrays["out"]
    assert out.shape == ref.shape and out.dtype == ref.dtype
    err = rel_l2(out, ref)
    assert err <= tolerance_for(ref.dtype, params), f"{name}: {err:.3e}"
    if ref.dtype == np.float64 and not params.get("parity_tol"):
        assert err <= 1e-12, f"{name}: {err:.3e}"


@pytest.mark.parametrize("shape,dim,tile,m,iso,het,dtype,batch", cg_fixed_count_cases())
@pytest.mark.parametrize("fused", [True, False])
def test_cg_fixed_count_single_pass_and_two_pass(da, shape, dim, tile, m, iso, het, dtype, batch, fused):
    """CG(tol=0): the single-pass route (fused=True, K7) and the two-kernel route agree with the oracle."""
    out, ref = run_cg_fixed_pair(da, oracle, shape, dim, tile, m, iso, het, dtype, batch, fused=fused)
    assert rel_l2(out, ref) <= (1e-13 if dtype == np.float64 else 2e-6)


@pytest.mark.parametrize("shape,dim,tile,K,iso,dtype,batch", jacobi_blocked_cases())
def test_jacobi_sweeps_temporally_blocked(da, shape, dim, tile, K, iso, dtype, batch):
    """Jacobi(K): S sweeps per pass (K4xS) == one sweep per pass (same arithmetic per voxel; bitwise in the
    CPU thread-stepper, to an ulp here should nvcc contract the two kernels differently), and both agree
    with the oracle."""
    out, one, ref = run_jacobi_blocked_pair(da, oracle, shape, dim, tile, K, iso, dtype, batch)
    assert rel_l2(out, one) <= (1e-15 if dtype == np.float64 else 1e-6)
    assert rel_l2(out, ref) <= (1e-13 if dtype == np.float64 else 2e-6)


@pytest.mark.parametrize("shape,tile,nit,iso,het,dtype,batch", jacobi1_pair_cases())
def test_jacobi1_two_iterations_per_launch(da, shape, tile, nit, iso, het, dtype, batch):
    """K1x2 == K1 applied twice (to the last bit in the CPU thread-stepper, which compiles without FMA
    contraction; nvcc contracts the two kernels' expressions differently, so here to a few ulp); both agree
    with the oracle, and the per-iteration increments of the paired launches are the oracle's."""
    out, one, ref, info, hist = run_jacobi1_pair(da, oracle, shape, tile, nit, iso, het, dtype, batch)
    assert rel_l2(out, one) <= (1e-14 if dtype == np.float64 else 1e-6)
    assert rel_l2(out, ref) <= (1e-13 if dtype == np.float64 else 2e-6)
    assert list(info.iters) == [nit] * max(1, int(np.prod(batch, dtype=int)))
    for b, h in enumerate(hist if batch else [hist]):
        np.testing.assert_allclose(info.history[: len(h), b], h, rtol=(1e-9 if dtype == np.float64 else 1e-3))


@pytest.mark.parametrize("shape,dim,tile,solver,iso", adaptive_cases())
def test_adaptive_schedule(da, shape, dim, tile, solver, iso):
    for tl in (tile, (0, 0, 0)):
        out, ref = run_adaptive_pair(da, oracle, shape, dim, tl, solver, iso)
        assert rel_l2(out, ref) <= 1e-11


@pytest.mark.parametrize("shape,dim,tile,iso,solver,hetero,dtype", list(small_grid()))
def test_tile_seams(da, shape, dim, tile, iso, solver, hetero, dtype):
    out, ref = run_pair(da, oracle, shape, dim, iso, solver, dtype=dtype, hetero=hetero, tile=tile)
    assert rel_l2(out, ref) <= (1e-12 if dtype == np.float64 else 2e-5)


MEDIUM = [
    ((301, 517), 2, False, "j1", False, np.float64, 25),
    ((301, 517), 2, True, "j1", True, np.float64, 25),
    ((301, 517), 2, True, "cg10", False, np.float64, 12),
    ((300, 516), 2, False, "j20", True, np.float64, 8),
    ((301, 517), 2, False, "cgtol", False, np.float64, 8),
    ((301, 517), 2, False, "jtol", False, np.float64, 8),
    ((301, 517), 2, False, "j1", False, np.float32, 25),
    ((301, 517), 2, True, "cg4", True, np.float32, 10),
    ((41, 70, 93), 3, False, "j1", False, np.float64, 20),
    ((41, 70, 93), 3, True, "j1", True, np.float64, 20),
    ((41, 70, 93), 3, True, "cg10", False, np.float64, 8),
    ((40, 64, 96), 3, False, "j3", False, np.float64, 10),
    ((41, 70, 93), 3, False, "j1", False, np.float32, 20),
    ((41, 70, 93), 3, True, "cg4", False, np.float32, 8),
]


@pytest.mark.parametrize("shape,dim,iso,solver,hetero,dtype,nit", MEDIUM)
def test_default_tiles_medium_shapes(da, shape, dim, iso, solver, hetero, dtype, nit):
    """Default launch shapes (several tiles and chunks), odd extents, full iteration counts."""
    out, ref = run_pair(da, oracle, shape, dim, iso, solver, dtype=dtype, hetero=hetero, nit=nit)
    assert rel_l2(out, ref) <= (1e-12 if dtype == np.float64 else 1e-5)


def test_batched_channels(da):
    out, ref = run_pair(da, oracle, (120, 260), 2, True, "cg4", batch=(3,))
    assert rel_l2(out, ref) <= 1e-12
    out, ref = run_pair(da, oracle, (120, 260), 2, False, "j1", batch=(5,), hetero=True)
    assert rel_l2(out, ref) <= 1e-12
    out, ref = run_pair(da, oracle, (20, 30, 40), 3, False, "j1", batch=(2,))
    assert rel_l2(out, ref) <= 1e-12


def test_fused_equals_building_blocks(da):
    from parity_cases import make_inputs

    img, kw = make_inputs((257, 400), hetero=True)
    a = da.split_bregman_tvd(img, max_num_iter=30, isotropic=True, **kw)
    b = da.split_bregman_tvd(img, max_num_iter=30, isotropic=True, fused=False, **kw)
    assert rel_l2(a, b) <= 1e-13


def test_eps_stops_each_problem_at_the_reference_iteration(da):
    from golden.generate_golden import phantom

    img = np.stack([phantom((24, 40), 3), 0.3 * phantom((24, 40), 4) + 0.3], axis=-1)
    hist = []
    ref = oracle.split_bregman_tvd(img, mu=0.1, ell=1.0, max_num_iter=60, eps=5e-3, history=hist,
                                   solver=oracle.CG(maxiter=3, tol=0.0))
    out, info = da.split_bregman_tvd(img, mu=0.1, ell=1.0, max_num_iter=60, eps=5e-3, return_info=True,
                                     solver=da.CG(maxiter=3, tol=0.0))
    assert [len(h) for h in hist] == list(info.iters)
    assert rel_l2(out, ref) <= 1e-12
    assert info.launched < 60  # the host stopped enqueueing after the device flag fired


def test_deterministic(da):
    from parity_cases import make_inputs

    img, kw = make_inputs((300, 500))
    a = da.split_bregman_tvd(img, max_num_iter=20, isotropic=True, solver=da.CG(maxiter=5, tol=0.0), **kw)
    b = da.split_bregman_tvd(img, max_num_iter=20, isotropic=True, solver=da.CG(maxiter=5, tol=0.0), **kw)
    np.testing.assert_array_equal(a, b)


def test_image_in_out_and_entry_point(da):
    from parity_cases import make_inputs

    img, _ = make_inputs((64, 80))
    im = da.Image(img.copy(), space_dim=2)
    out = da.tvd(im, method="heterogeneous bregman", weight=0.1, omega=1.0, max_num_iter=20, eps=1e-3)
    ref = oracle.tvd(img, method="heterogeneous bregman", weight=0.1, omega=1.0, max_num_iter=20, eps=1e-3)
    assert isinstance(out, da.Image) and rel_l2(out.img, ref) <= 1e-12


def test_standalone_solvers(da):
    rng = np.random.default_rng(2)
    x0, rhs = rng.random((90, 140)), rng.standard_normal((90, 140))
    for mk in (lambda m: m.Jacobi(maxiter=4, dim=2, mass_coeff=1.3, diffusion_coeff=0.7),
               lambda m: m.CG(maxiter=6, tol=0.0, dim=2, mass_coeff=1.3, diffusion_coeff=0.7)):
        s_ref = mk(oracle)
        s_ref.reset()
        assert rel_l2(mk(da)(x0=x0, rhs=rhs), s_ref(x0=x0, rhs=rhs)) < 1e-12


@pytest.mark.parametrize("name", ["vol_jac_tol", "vol_cg", "rgb_jac5", "rgb_jac_tol", "vol_mg", "vol_mg_hetero"])
def test_h1_regularization_golden(da, name):
    import h1_cases

    out, ref = h1_cases.run(da, h1_cases.load(), name)
    assert out.shape == ref.shape and rel_l2(out, ref) <= 1e-12


def test_state_round_trip_and_device_resident_io(da):
    """Device-resident callers: CUDA tensors in -> CUDA tensors out, and the (u, d, b) state of
    one call warm-starts the next: 4 + 3 iterations == 7 iterations straight (oracle)."""
    import torch
    from golden.generate_golden import coeff_field, phantom

    img = phantom((120, 131, 2), 3)
    mu = coeff_field((120, 131), 8, 0.05, 0.1, 0.1)
    kw = dict(omega=1.0, ell=1.0, dim=2, isotropic=True)
    ref7 = oracle.split_bregman_tvd(img, mu=mu, max_num_iter=7, solver=oracle.CG(maxiter=4, tol=0.0), **kw)
    st = da.split_bregman_tvd(img, mu=mu, max_num_iter=4, solver=da.CG(maxiter=4, tol=0.0), return_state=True, **kw)
    out = da.split_bregman_tvd(img, mu=mu, max_num_iter=3, solver=da.CG(maxiter=4, tol=0.0), x0=st, **kw)
    assert rel_l2(out, ref7) <= 1e-12
    dev = torch.device("cuda", 0)
    ti, tmu = torch.from_numpy(img).to(dev), torch.from_numpy(mu).to(dev)
    tst = da.split_bregman_tvd(ti, mu=tmu, max_num_iter=4, solver=da.CG(maxiter=4, tol=0.0), return_state=True, **kw)
    assert all(isinstance(a, torch.Tensor) and a.is_cuda for a in tst)
    tout = da.split_bregman_tvd(ti, mu=tmu, max_num_iter=3, solver=da.CG(maxiter=4, tol=0.0), x0=tst, **kw)
    assert tout.is_cuda and np.array_equal(tout.cpu().numpy(), out)

from conftest import GOLDEN_DIR  # noqa: E402


@pytest.mark.parametrize("tag,dim", [("2d", 2), ("3d", 3)])
def test_mg_standalone_and_errors(da, tag, dim):
    from parity_cases import check_mg_standalone

    check_mg_standalone(da, tag, dim)


def test_mg_warm_start_and_state_round_trip(da):
    from parity_cases import check_mg_warm_start

    check_mg_warm_start(da, oracle)


from parity_cases import mg_hetero_cases, run_mg_hetero_pair  # noqa: E402


@pytest.mark.parametrize("shape,dim,depth,iso,dtype,batch,which", mg_hetero_cases())
def test_mg_per_voxel_coefficients(da, shape, dim, depth, iso, dtype, batch, which):
    out, ref = run_mg_hetero_pair(da, oracle, shape, dim, depth, iso, dtype, batch, which)
    assert rel_l2(out, ref) <= (1e-12 if dtype == np.float64 else 2e-5)


def test_mg_inner_solver_batched_and_stopping(da):
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


def test_series_sub_batches_overlapped(da):
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
def test_chambolle_matches_its_restatement(da, shape, dtype, weight, eps, nit):
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


def test_tvd_default_method_is_chambolle(da):
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


@pytest.mark.parametrize("case", fuzz_cases.draw(seed=12, n=80), ids=lambda c: f"{c['kind']}-{c['dim']}d-f{c['fused']}")
def test_random_problems_over_every_route(da, case):
    """Seeded random shapes / tiles / batches / dtypes / coefficient maps / solvers / kernel routes vs the oracle."""
    out, ref = fuzz_cases.run(da, oracle, case)
    assert rel_l2(out, ref) <= (1e-12 if case["dtype"] == np.float64 else 1e-5)  # north_star's bars are 1e-10 / 1e-5


def test_single_pass_cg_predicted_rho_diagnostic(da):
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


@pytest.mark.parametrize("shape,batch,tile,dtype,nit", tma_cases() + [((70, 90, 200), (), (0, 0, 0), np.float64, 8),
                                                                         ((130, 61, 257), (), (0, 0, 0), np.float32, 8)])
@pytest.mark.parametrize("iso", [False, True])
def test_tma_staged_fused_jacobi1_3d(da, shape, batch, tile, dtype, nit, iso):
    """KFusedJ1T on the B200: cp.async.bulk.tensor boxes + mbarrier ring against the oracle (and the march kernel's bar)."""
    out, ref, info, hist = run_tma_case(da, oracle, shape, batch, tile, dtype, nit, iso)
    assert rel_l2(out, ref) <= (1e-13 if dtype == np.float64 else 2e-6)
    nb = max(1, int(np.prod(batch, dtype=int)))
    assert list(info.iters) == [nit] * nb
    np.testing.assert_allclose(info.history[:nit].reshape(-1), np.array(hist).reshape(nb, nit).T.reshape(-1),
                               rtol=1e-9 if dtype == np.float64 else 1e-3)


@pytest.mark.parametrize("shape,batch,tile,dtype,nit", tma2d_cases() + [((447, 795), (), (0, 0, 0), np.float64, 12),
                                                                           ((1788, 3180), (), (0, 0, 0), np.float64, 4),
                                                                           ((300, 1000), (2,), (96, 0, 0), np.float64, 6),
                                                                           ((500, 700), (), (0, 0, 0), np.float32, 8)])
def test_tma_staged_fused_jacobi1_2d(da, shape, batch, tile, dtype, nit):
    """KFusedJ1T<T, false, 2> on the B200 (fused = 4): row boxes through cp.async.bulk.tensor + mbarrier ring against the
    oracle and the march kernel."""
    out, ref, info, hist, march = run_tma2d_case(da, oracle, shape, batch, tile, dtype, nit)
    assert rel_l2(out, ref) <= (1e-13 if dtype == np.float64 else 2e-6)
    assert rel_l2(out, march) <= (1e-14 if dtype == np.float64 else 1e-6)
    nb = max(1, int(np.prod(batch, dtype=int)))
    assert list(info.iters) == [nit] * nb
    np.testing.assert_allclose(info.history[:nit].reshape(-1), np.array(hist).reshape(nb, nit).T.reshape(-1),
                               rtol=1e-9 if dtype == np.float64 else 1e-3)


def test_series_streaming_pipeline(da):
    """tvd_series_stream: frames from a callable / an iterator through the staging rings == frame-by-frame oracle; partial
    last sub-batch, colour channels, shared per-pixel maps, sink delivery in order."""
    from golden.generate_golden import coeff_field, phantom

    shape = (180, 250)
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
    # frames the caller keeps in pinned host memory are DMA'd as they are (no staging copy): same result
    import torch

    pinned = [torch.from_numpy(fr).pin_memory().numpy() for fr in frames]
    got_p, _ = da.tvd_series_stream(lambda k: pinned[k], n_frames=7, dim=2, batch=3, mu=mu, omega=om, ell=1.0, max_num_iter=6,
                                    solver=da.Jacobi(maxiter=1))
    assert all(np.array_equal(a, b) for a, b in zip(got_p, got))


@pytest.mark.parametrize("solver", ["cg", "j3"])
@pytest.mark.parametrize("shape,tile,dtype", stencil_cases())
def test_tma_staged_stencil_3d(da, shape, tile, dtype, solver):
    """KStencilT (3-D cg_pq / Jacobi sweep with scalar coefficients) against the oracle."""
    out, ref = run_stencil_case(da, oracle, shape, tile, dtype, solver)
    assert rel_l2(out, ref) <= (1e-13 if dtype == np.float64 else 2e-6)
