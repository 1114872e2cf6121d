"""CPU: the Python boundary -- argument plumbing of TVD/tvd (reference tvd.py:23-63), the
Image in/out contract, error behaviour, solver descriptors.  Kernels run in the emulator."""

import numpy as np
import pytest

import oracle
from conftest import rel_l2
from emu.harness import emulated

import darsia_b200 as da


@pytest.fixture(scope="module", autouse=True)
def _emu():
    with emulated():
        yield


@pytest.fixture
def img():
    return np.clip(0.5 + 0.2 * np.random.default_rng(1).standard_normal((21, 34)), 0, 1)


def test_tvd_defaults_and_key_prefix(img):
    # reference tests/unit/test_restoration_config.py:39-98 pins these defaults on the config side
    op = da.TVD(method="heterogeneous bregman")
    assert (op.weight, op.max_num_iter, op.eps, op.omega, op.regularization) == (0.1, 200, 2e-4, 1, 1.0)
    # weight/max_num_iter/eps/method honour the key prefix, omega/regularization do not (Q15)
    cfg = {"restoration method": "Heterogeneous Bregman", "restoration weight": 0.3, "restoration max_num_iter": 7,
           "restoration eps": None, "omega": 0.5, "regularization": 0.7, "dim": 2}
    op = da.TVD(key="restoration ", **cfg)
    assert op.method == "heterogeneous bregman" and op.weight == 0.3 and op.max_num_iter == 7 and op.eps is None
    assert op.omega == 0.5 and op.regularization == 0.7 and op.kwargs == {"dim": 2}  # Q1: popped
    out = op(img)
    ref = oracle.TVD(key="restoration ", **cfg)(img)
    assert rel_l2(out, ref) < 1e-12


def test_call_has_two_positional_args():
    # CombinedModel checks model.__call__.__code__.co_argcount == 2 (combinedmodel.py:47-48)
    assert da.TVD.__call__.__code__.co_argcount == 2


def test_workflow_style_weight_mask(img):
    # reference tests/unit/test_restoration_workflow.py:62-98: weight * mask, forced heterogeneous bregman
    mask = np.ones(img.shape)
    mask[:, :10] = 0.0
    weight = 0.1 * mask
    out = da.tvd(img, method="heterogeneous bregman", weight=weight, omega=1.0, regularization=1.0,
                 max_num_iter=10, eps=None)
    ref = oracle.tvd(img, method="heterogeneous bregman", weight=weight, omega=1.0, regularization=1.0,
                     max_num_iter=10, eps=None)
    assert rel_l2(out, ref) < 1e-12


def test_image_in_image_out(img):
    im = da.Image(img.copy(), space_dim=2, name="x")
    out = da.tvd(im, method="heterogeneous bregman", weight=0.1, max_num_iter=3, eps=None)
    assert isinstance(out, da.Image) and out is not im and out.metadata() == im.metadata()
    np.testing.assert_array_equal(im.img, img)  # input untouched
    ref = oracle.tvd(img, method="heterogeneous bregman", weight=0.1, max_num_iter=3, eps=None)
    assert rel_l2(out.img, ref) < 1e-12


def test_errors(img):
    with pytest.raises(TypeError):
        da.tvd([1, 2, 3], method="heterogeneous bregman")
    with pytest.raises(ValueError):
        da.tvd(img, method="no such method")
    with pytest.raises(NotImplementedError):
        da.tvd(img, method="anisotropic bregman")  # scikit-image's sequential Gauss-Seidel scheme
    with pytest.raises(TypeError):
        da.split_bregman_tvd((img * 255).astype(np.int32))  # convert_dtype returns uint8 / uint16 / floats only
    with pytest.raises(TypeError):
        da.split_bregman_tvd(img, solver=da.CG(maxiter=3))  # tol=None is not a tolerance (Q4)
    with pytest.raises(NotImplementedError):
        da.split_bregman_tvd(img, solver=da.Solver())
    with pytest.raises(ValueError):
        da.split_bregman_tvd(img, mu=np.zeros(img.shape))  # ell = 2*mu = 0 (Q12)
    with pytest.raises(AssertionError):
        da.split_bregman_tvd(img, mu=np.ones((3, 3)))
    with pytest.raises(TypeError):
        da.split_bregman_tvd(img, adaptive=True)  # a schedule is a callable iter -> bool


def test_dtype_and_precision(img):
    out32 = da.split_bregman_tvd(img.astype(np.float32), mu=0.1, ell=1.0, max_num_iter=5)
    assert out32.dtype == np.float32
    out = da.split_bregman_tvd(img, mu=0.1, ell=1.0, max_num_iter=5, precision="fp32")
    assert out.dtype == np.float64
    ref = oracle.split_bregman_tvd(img, mu=0.1, ell=1.0, max_num_iter=5)
    assert rel_l2(out, ref) < 1e-5 and rel_l2(out32, ref) < 1e-5


def test_verbose_prints_reference_lines(img, capsys):
    da.split_bregman_tvd(img, mu=0.1, ell=1.0, max_num_iter=4, verbose=2)
    lines = capsys.readouterr().out.strip().splitlines()
    assert lines[0].startswith("The energy functional starts at ")
    assert [ln.split(" - ")[0] for ln in lines[1:]] == ["Split Bregman iteration 0", "Split Bregman iteration 2"]
    # energy printed at iteration k is the functional of the iterate BEFORE the update (:162-167)
    e0 = float(lines[0].rsplit(" ", 1)[1])
    mu, om = 0.1, 1.0
    expect = 0.0  # f - f = 0, plus TV of f
    for ax in range(2):
        expect += np.sum(np.abs(mu * oracle.backward_diff(img, ax, 2)))
    assert abs(e0 - expect) < 1e-9 * max(1.0, expect)
    da.split_bregman_tvd(img, mu=0.1, ell=1.0, max_num_iter=2, verbose=0)  # Q11: silent, no ZeroDivisionError
    assert capsys.readouterr().out == ""


def test_standalone_solvers(img):
    rhs = np.random.default_rng(2).standard_normal(img.shape)
    for mk in (lambda m: m.Jacobi(maxiter=4, dim=2, mass_coeff=1.3, diffusion_coeff=0.7),
               lambda m: m.Jacobi(maxiter=30, tol=1e-2, dim=2, mass_coeff=1.3, diffusion_coeff=0.7),
               lambda m: m.CG(maxiter=6, tol=0.0, dim=2, mass_coeff=1.3, diffusion_coeff=0.7),
               lambda m: m.CG(maxiter=100, tol=1e-6, dim=2, mass_coeff=1.3, diffusion_coeff=0.7)):
        s_ref = mk(oracle)
        s_ref.reset()
        assert rel_l2(mk(da)(x0=img, rhs=rhs), s_ref(x0=img, rhs=rhs)) < 1e-12


def test_reference_example_script_runs(tmp_path, capsys):
    """examples/regularization.py is the reference's example with ``darsia`` replaced by ``darsia_b200``
    (heterogeneous weight, Jacobi(20), 3-D; then H1 regularisation)."""
    import os
    import runpy
    import sys

    from conftest import ROOT
    from golden.generate_golden import phantom

    path = tmp_path / "vol.npy"
    np.save(path, phantom((12, 20, 28), 3))
    argv = sys.argv
    sys.argv = ["examples/regularization.py", str(path)]
    try:
        runpy.run_path(os.path.join(ROOT, "examples", "regularization.py"), run_name="__main__")
    finally:
        sys.argv = argv
    out = capsys.readouterr().out
    assert "TVD" in out and "H1" in out


@pytest.mark.parametrize("dtype,top", [(np.uint8, 255), (np.uint16, 65535)])
def test_integer_images_follow_img_as_float_and_convert_dtype(img, dtype, top):
    """uint8 / uint16 in (what tvd(jpg_array) gets): skimage.img_as_float on the way in (split_bregman_tvd.py:94,99),
    convert_dtype = img_as_ubyte / img_as_uint on the way out (:177, utils/dtype.py:9-30); the solve runs in fp64."""
    raw = np.rint(img * top).astype(dtype)
    out = da.split_bregman_tvd(raw, mu=0.1, ell=1.0, max_num_iter=6)
    assert out.dtype == dtype and out.shape == raw.shape
    ref = oracle.split_bregman_tvd(raw.astype(np.float64) / top, mu=0.1, ell=1.0, max_num_iter=6)
    want = np.clip(np.rint(ref * top), 0, top).astype(dtype)
    assert int(np.max(np.abs(out.astype(np.int64) - want.astype(np.int64)))) <= 1  # a rounding tie at most
    assert np.mean(out != want) < 1e-3
    out2 = da.tvd(raw, method="heterogeneous bregman", weight=0.1, max_num_iter=3, eps=None)
    assert out2.dtype == dtype


def test_zero_image_with_warm_start_and_cg_returns_zero(img):
    """scipy cg returns b when ||b|| = 0 (cg.py:50-56 -> scipy): with a zero image every right-hand side is zero, so the
    iterate becomes zero whatever the warm start was."""
    z = np.zeros_like(img)
    x0 = (img.copy(), np.zeros(img.shape + (2,)), np.zeros(img.shape + (2,)))
    ref = oracle.split_bregman_tvd(z, mu=0.1, ell=1.0, max_num_iter=2, solver=oracle.CG(maxiter=3, tol=0.0), x0=x0)
    with np.errstate(all="ignore"):
        import warnings

        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            out = da.split_bregman_tvd(z, mu=0.1, ell=1.0, max_num_iter=2, solver=da.CG(maxiter=3, tol=0.0), x0=x0)
    assert np.array_equal(ref, np.zeros_like(ref)) and np.array_equal(out, ref)
