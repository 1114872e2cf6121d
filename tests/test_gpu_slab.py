"""GPU (one device is enough): the slab-decomposed solve through ``tvdk_solve_slab``.

All slabs of the decomposition live in THIS process and on the same device, driven in lock step on one
stream: the slab-geometry paths of the kernels (``g0_off``, ``z_begin``/``z_end``), the peer pushes of halo
planes and partial sums, and the ordered cross-rank sums all run on the GPU the driver tests on, without any
kernel waiting on a concurrently running one (every flag is already raised, by stream order, when it is read).
The decomposed result must equal the single-volume solve of the same library (<= 1e-12: only the order of the
partial sums differs) and the oracle (<= 1e-10 / 1e-5, BASELINE north_star)."""

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

SOLVERS = {
    "j1": lambda m: m.Jacobi(maxiter=1),
    "j3": lambda m: m.Jacobi(maxiter=3),
    "jtol": lambda m: m.Jacobi(maxiter=12, tol=2e-2),
    "cg5": lambda m: m.CG(maxiter=5, tol=0.0),
    "cgtol": lambda m: m.CG(maxiter=30, tol=1e-3),
}

CASES = [
    # world, shape, dtype, iso, solver, nit, eps, hetero
    (2, (40, 36, 70), "float64", False, "j1", 12, None, False),
    (3, (41, 36, 70), "float64", True, "j1", 10, None, True),
    (2, (40, 36, 70), "float32", False, "j1", 12, None, False),
    (3, (40, 36, 70), "float64", True, "cg5", 6, None, False),
    (2, (40, 36, 70), "float64", False, "cgtol", 4, None, False),
    (2, (40, 36, 70), "float64", False, "j3", 5, None, False),
    (2, (40, 36, 70), "float64", False, "jtol", 5, None, False),
    (4, (64, 24, 40), "float64", False, "j1", 60, 2e-2, False),
    (2, (300, 517), "float64", True, "cg5", 5, None, False),
    (8, (96, 20, 36), "float64", True, "cg5", 4, None, False),
]


@pytest.mark.parametrize("world,shape,dtype,iso,solver,nit,eps,hetero", CASES)
def test_slabs_on_one_device_equal_single_volume_and_oracle(world, shape, dtype, iso, solver, nit, eps, hetero):
    import darsia_b200 as da
    from darsia_b200.slab_p2p import SlabGroup

    vol = phantom(shape, 3, np.dtype(dtype))
    kw = dict(mu=0.1, omega=1.0, ell=1.0)
    if hetero:
        kw.update(mu=coeff_field(shape, 5, 0.05, 0.1, 0.1).astype(vol.dtype), omega=coeff_field(shape, 6, 0.5, 2.0).astype(vol.dtype))
    dim = len(shape)
    grp = SlabGroup.local(shape, dim, vol.dtype, kw["mu"], kw["omega"], kw["ell"], iso, SOLVERS[solver](da), eps=eps, world=world)
    grp.setup([np.ascontiguousarray(vol[slice(*da.slab_range(shape[0], r, world))]) for r in range(world)])
    grp.solve(nit)
    grp.check()
    out = np.concatenate(grp.results(), axis=0)
    hists = [grp.history(i) for i in range(world)]
    single, info = da.split_bregman_tvd(vol, dim=dim, max_num_iter=nit, isotropic=iso, solver=SOLVERS[solver](da), eps=eps,
                                        return_info=True, **kw)
    hist = []
    ref = oracle.split_bregman_tvd(vol, dim=dim, max_num_iter=nit, isotropic=iso, solver=SOLVERS[solver](oracle), eps=eps,
                                   history=hist, **kw)
    f64 = vol.dtype == np.float64
    assert rel_l2(out, single) <= (1e-12 if f64 else 2e-6)
    assert rel_l2(out, ref) <= (1e-10 if f64 else 1e-5)
    assert len(hists[0]) == len(hist) == int(info.iters[0])
    np.testing.assert_allclose(hists[0], hist, rtol=1e-6 if f64 else 1e-3)
    for h in hists[1:]:
        assert h == hists[0]


def test_second_solve_on_the_same_arenas_is_bitwise_identical():
    import darsia_b200 as da
    from darsia_b200.slab_p2p import SlabGroup

    shape = (48, 40, 72)
    vol = phantom(shape, 9)
    grp = SlabGroup.local(shape, 3, vol.dtype, 0.1, 1.0, 1.0, True, da.CG(maxiter=4, tol=0.0), world=3)
    grp.setup([np.ascontiguousarray(vol[slice(*da.slab_range(shape[0], r, 3))]) for r in range(3)])
    grp.solve(5)
    a = np.concatenate(grp.results(), axis=0)
    grp.reset()
    grp.solve(5)
    grp.check()
    b = np.concatenate(grp.results(), axis=0)
    assert np.array_equal(a, b)
