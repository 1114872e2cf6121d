"""GPU (B200): parity at BASELINE.json's full sizes.  Full-size, full-iteration oracle runs are
infeasible on the CPU, so: a few outer iterations against the oracle at full 2-D
size, plus size-independent properties at full size and iteration count (a constant image is a
fixed point; batched channels == channels alone, bitwise; fp32 mode within 1e-5 of fp64; the
fused and the building-block routes agree)."""

import numpy as np
import pytest

import oracle
from conftest import rel_l2

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def da():
    import darsia_b200

    return darsia_b200


def _phantom(shape, seed, dtype=np.float64):
    import bench

    return bench.phantom_np(shape, seed, dtype)


def test_cfg1_full_size_vs_oracle(da):
    """examples/regularization.py-sized 2-D image (1788x3180), anisotropic, Jacobi."""
    img = _phantom((1788, 3180), 1)
    for solver, nit in ((lambda m: m.Jacobi(maxiter=1), 6), (lambda m: m.Jacobi(maxiter=20), 2)):
        out = da.tvd(img, method="heterogeneous bregman", weight=0.1, omega=1.0, dim=2, isotropic=False,
                     max_num_iter=nit, eps=None, solver=solver(da))
        ref = oracle.tvd(img, method="heterogeneous bregman", weight=0.1, omega=1.0, dim=2, isotropic=False,
                         max_num_iter=nit, eps=None, solver=solver(oracle))
        assert rel_l2(out, ref) <= 1e-10


def test_cfg2_full_size(da):
    """4000x7000 3-channel, isotropic, CG: one channel vs the oracle for 2 outer iterations, and the
    3-channel batch vs the channels alone (bitwise) at the same settings."""
    chans = [_phantom((4000, 7000), 2 + c) for c in range(3)]
    img = np.stack(chans, axis=-1)
    kw = dict(mu=0.1, omega=1.0, ell=1.0, dim=2, isotropic=True, max_num_iter=2)
    out = da.split_bregman_tvd(img, solver=da.CG(maxiter=3, tol=0.0), **kw)
    ref0 = oracle.split_bregman_tvd(chans[0], solver=oracle.CG(maxiter=3, tol=0.0), **kw)
    assert rel_l2(out[..., 0], ref0) <= 1e-10
    one = da.split_bregman_tvd(chans[1], solver=da.CG(maxiter=3, tol=0.0), **kw)
    np.testing.assert_array_equal(out[..., 1], one)


def test_cfg2_constant_image_is_fixed_point_100_iterations(da):
    """A constant image is a fixed point (rhs = omega*f, Laplacian of a constant is 0).  With CG the
    tolerance must be positive: for tol = 0 SciPy's recurrence itself divides 0/0 on a zero residual
    (the reference returns NaN there, and so do the oracle and this implementation)."""
    img = np.full((4000, 7000), 0.37)
    out, info = da.split_bregman_tvd(img, mu=0.1, omega=1.0, ell=1.0, dim=2, isotropic=True, max_num_iter=100,
                                     solver=da.CG(maxiter=10, tol=1e-8), return_info=True)
    assert int(info.iters[0]) == 100
    assert rel_l2(out, img) <= 1e-12
    out = da.split_bregman_tvd(img, mu=0.1, omega=1.0, ell=1.0, dim=2, isotropic=True, max_num_iter=100)
    assert rel_l2(out, img) <= 1e-12


def test_full_size_jacobi1_pairs_and_blocked_sweeps(da):
    """4000x7000 anisotropic fp64 with scalar coefficients: Jacobi(1) takes the two-iterations-per-launch kernel
    (K1x2) by default here, Jacobi(9) the temporally blocked sweeps (K4x4): both against the oracle, and against
    the one-step-per-launch route at a higher iteration count."""
    img = _phantom((4000, 7000), 7)
    kw = dict(mu=0.1, omega=1.0, ell=1.0, dim=2, isotropic=False)
    out = da.split_bregman_tvd(img, max_num_iter=5, **kw)  # 2 pairs + 1 single
    ref = oracle.split_bregman_tvd(img, max_num_iter=5, **kw)
    assert rel_l2(out, ref) <= 1e-10
    a = da.split_bregman_tvd(img, max_num_iter=40, **kw)
    b = da.split_bregman_tvd(img, max_num_iter=40, fused=3, **kw)  # 3: one iteration per launch
    assert rel_l2(a, b) <= 1e-13
    out = da.split_bregman_tvd(img, max_num_iter=2, solver=da.Jacobi(maxiter=9), **kw)
    ref = oracle.split_bregman_tvd(img, max_num_iter=2, solver=oracle.Jacobi(maxiter=9), **kw)
    assert rel_l2(out, ref) <= 1e-10


def test_cfg3_hetero_frames_full_size(da):
    """Two 4000x7000 frames with per-pixel mu (10 % zero mask) and omega, fused Jacobi(1): frame 0 vs the
    oracle after 4 iterations; fused route == building-block route after 30."""
    import bench

    w = dict(shape=(4000, 7000), dtype="float64")
    mu, om = bench.hetero_maps(w, 2)
    frames = np.stack([_phantom((4000, 7000), 3 + t) for t in range(2)], axis=-1)
    kw = dict(mu=mu, omega=om, ell=1.0, dim=2, isotropic=False)
    out = da.split_bregman_tvd(frames, max_num_iter=4, **kw)
    ref = oracle.split_bregman_tvd(frames[..., 0].copy(), max_num_iter=4, **kw)
    assert rel_l2(out[..., 0], ref) <= 1e-10
    a = da.split_bregman_tvd(frames, max_num_iter=30, **kw)
    b = da.split_bregman_tvd(frames, max_num_iter=30, fused=False, **kw)
    assert rel_l2(a, b) <= 1e-12


def test_3d_volume_vs_oracle_and_fp32_mode(da):
    """CT-like volume: 160^3 vs the oracle (isotropic CG and anisotropic Jacobi), then a 320x320x256
    volume for 50 iterations in fp32 mode against fp64 (cfg5's tolerance check, <= 1e-5)."""
    vol = _phantom((160, 160, 160), 4)
    out = da.split_bregman_tvd(vol, mu=0.1, omega=1.0, ell=1.0, dim=3, isotropic=True, max_num_iter=3,
                               solver=da.CG(maxiter=10, tol=0.0))
    ref = oracle.split_bregman_tvd(vol, mu=0.1, omega=1.0, ell=1.0, dim=3, isotropic=True, max_num_iter=3,
                                   solver=oracle.CG(maxiter=10, tol=0.0))
    assert rel_l2(out, ref) <= 1e-10
    out = da.split_bregman_tvd(vol, mu=0.1, omega=1.0, ell=1.0, dim=3, isotropic=False, max_num_iter=10)
    ref = oracle.split_bregman_tvd(vol, mu=0.1, omega=1.0, ell=1.0, dim=3, isotropic=False, max_num_iter=10)
    assert rel_l2(out, ref) <= 1e-10
    big = _phantom((320, 320, 256), 5)
    o64 = da.split_bregman_tvd(big, mu=0.1, omega=1.0, ell=1.0, dim=3, isotropic=False, max_num_iter=50)
    o32 = da.split_bregman_tvd(big, mu=0.1, omega=1.0, ell=1.0, dim=3, isotropic=False, max_num_iter=50,
                               precision="fp32")
    assert o32.dtype == np.float64 and rel_l2(o32, o64) <= 1e-5
    n32 = da.split_bregman_tvd(big.astype(np.float32), mu=0.1, omega=1.0, ell=1.0, dim=3, isotropic=False,
                               max_num_iter=50)
    assert n32.dtype == np.float32 and rel_l2(n32, o64) <= 1e-5
