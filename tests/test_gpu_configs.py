"""GPU: the BASELINE configs at their FULL iteration counts on geometrically similar reduced shapes (SURVEY 8d parity
protocol (i): odd and even extents, several tiles per axis), against the oracle:

  cfg2  isotropic, CG(maxiter=10, tol=0), 100 outer iterations, 3 channels    (single-pass CG route, predicted rho)
  cfg3  anisotropic, per-pixel mu / omega, Jacobi(1), 100 outer iterations, frames
  cfg4  3-D isotropic, CG(10), 20 outer iterations                             (two-kernel CG route)
  cfg5  3-D anisotropic, Jacobi(1), 50 outer iterations, fp64 and fp32
  cfg1  examples/regularization.py: Jacobi(20), 30 outer iterations, omega = 0.1

Bar (BASELINE north_star): relative L2 <= 1e-10 in fp64, <= 1e-5 in fp32.  Also here: the verbose energy functional
(split_bregman_tvd.py:71-77, kernel KEnergy) against the oracle's on the GPU."""

import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

import oracle  # noqa: E402
from conftest import rel_l2  # noqa: E402
from golden.generate_golden import coeff_field, phantom  # noqa: E402


def _both(img, mk, **kw):
    import darsia_b200 as da

    out, info = da.split_bregman_tvd(img, solver=mk(da), return_info=True, **kw)
    ref = oracle.split_bregman_tvd(img, solver=mk(oracle), **kw)
    return out, ref, info


def test_cfg2_iso_cg10_100_outer_three_channels():
    img = phantom((131, 771, 3), 21)
    out, ref, info = _both(img, lambda m: m.CG(maxiter=10, tol=0.0), mu=0.1, omega=1.0, ell=1.0, dim=2, isotropic=True,
                           max_num_iter=100)
    assert rel_l2(out, ref) <= 1e-10
    assert list(info.iters) == [100, 100, 100]
    # the single-pass route predicts rho for beta: the prediction must stay at rounding level over 1000 CG iterations
    assert float(np.max(info.cg_pred_dev)) < 1e-9


def test_cfg3_hetero_jacobi1_100_outer_frames():
    shape = (187, 771)
    frames = np.stack([phantom(shape, 30 + k) for k in range(2)], axis=-1)
    mu = 0.1 * coeff_field(shape, 5, 0.5, 1.0, 0.1)  # porosity-like weight with a 10 % zero mask
    om = coeff_field(shape, 6, 0.5, 2.0)
    out, ref, info = _both(frames, lambda m: m.Jacobi(maxiter=1), mu=mu, omega=om, ell=1.0, dim=2, isotropic=False,
                           max_num_iter=100)
    assert rel_l2(out, ref) <= 1e-10
    assert list(info.iters) == [100, 100]


def test_cfg4_3d_iso_cg10_20_outer():
    vol = phantom((64, 72, 136), 41)
    out, ref, _ = _both(vol, lambda m: m.CG(maxiter=10, tol=0.0), mu=0.1, omega=1.0, ell=1.0, dim=3, isotropic=True,
                        max_num_iter=20)
    assert rel_l2(out, ref) <= 1e-10


@pytest.mark.parametrize("dtype,bar", [(np.float64, 1e-10), (np.float32, 1e-5)])
def test_cfg5_3d_aniso_jacobi1_50_outer(dtype, bar):
    vol64 = phantom((96, 100, 136), 51)
    vol = vol64.astype(dtype)
    out, ref, _ = _both(vol, lambda m: m.Jacobi(maxiter=1), mu=0.1, omega=1.0, ell=1.0, dim=3, isotropic=False, max_num_iter=50)
    assert out.dtype == dtype
    assert rel_l2(out, ref) <= bar
    if dtype == np.float32:  # the fp32 mode against the fp64 reference result (BASELINE configs[4])
        ref64 = oracle.split_bregman_tvd(vol64, solver=oracle.Jacobi(maxiter=1), mu=0.1, omega=1.0, ell=1.0, dim=3, isotropic=False,
                                         max_num_iter=50)
        assert rel_l2(out, ref64) <= 1e-5


def test_cfg1_example_jacobi20_30_outer():
    img = phantom((223, 397), 61)
    out, ref, _ = _both(img, lambda m: m.Jacobi(maxiter=20), mu=0.1, omega=0.1, ell=1.0, dim=2, isotropic=False, max_num_iter=30)
    assert rel_l2(out, ref) <= 1e-10


@pytest.mark.parametrize("shape,dim,hetero", [((67, 130), 2, False), ((67, 130), 2, True), ((21, 30, 44), 3, True),
                                              ((40, 90, 2), 2, False)])
def test_verbose_energy_on_the_gpu_equals_the_reference_functional(shape, dim, hetero):
    """info.energy[k] is _functional(iterate entering iteration k) (split_bregman_tvd.py:162-167 prints it)."""
    import darsia_b200 as da

    img = phantom(shape, 71)
    spatial = shape[:dim]
    mu = 0.1 * coeff_field(spatial, 5, 0.5, 1.0, 0.1) if hetero else 0.1
    om = coeff_field(spatial, 6, 0.5, 2.0) if hetero else 1.3
    n = 6
    kw = dict(mu=mu, omega=om, ell=1.0, dim=dim, isotropic=False)
    _, info = da.split_bregman_tvd(img, solver=da.Jacobi(maxiter=2), max_num_iter=n, verbose=1, return_info=True, **kw)
    import contextlib
    import io

    nb = int(np.prod(shape[dim:]))
    flat = img.reshape(spatial + (nb,))
    for k in range(n):
        with contextlib.redirect_stdout(io.StringIO()):
            it_k = oracle.split_bregman_tvd(img, solver=oracle.Jacobi(maxiter=2), max_num_iter=k, **kw).reshape(spatial + (nb,))
        for b in range(nb):
            want = oracle.energy_functional(it_k[..., b], flat[..., b], mu, om, dim)
            assert abs(info.energy[k, b] - want) <= 1e-11 * max(1.0, abs(want)), (k, b)
