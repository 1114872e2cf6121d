"""CPU: the C slab driver (``tvdk_solve_slab``, include/tvdk.h) with the kernels in the emulator.

Two ways of driving the same code:
  * lock step -- one host thread drives every slab (what the 1-GPU test in test_gpu_slab.py does on cuda:0);
  * one thread per slab -- every slab has its own driver thread, like one process per GPU: the threads really
    wait on each other's mailbox flags (ctypes releases the GIL during the call).
The decomposed result must equal the oracle on the whole volume (same arithmetic per voxel; only the order of
the per-rank partial sums differs) and the increments history must match."""

import os
import sys
import threading

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import oracle  # noqa: E402
from conftest import rel_l2  # noqa: E402
from emu.harness import emulated  # noqa: E402
from golden.generate_golden import coeff_field, phantom  # noqa: E402

SOLVERS = {
    "j1": lambda m: m.Jacobi(maxiter=1),
    "j3": lambda m: m.Jacobi(maxiter=3),
    "jtol": lambda m: m.Jacobi(maxiter=12, tol=2e-2),
    "cg4": lambda m: m.CG(maxiter=4, tol=0.0),
    "cgtol": lambda m: m.CG(maxiter=30, tol=1e-3),
}

CASES = [
    # world, shape, dtype, iso, solver, nit, eps, hetero
    (2, (11, 9, 20), "float64", False, "j1", 6, None, False),
    (3, (13, 9, 20), "float64", True, "j1", 6, None, True),
    (2, (11, 9, 20), "float64", True, "cg4", 5, None, False),
    (3, (14, 7, 18), "float64", False, "cgtol", 4, None, False),
    (2, (11, 9, 20), "float64", False, "j3", 5, None, False),
    (2, (10, 8, 24), "float32", False, "j1", 8, None, False),
    (2, (40, 70), "float64", False, "j1", 6, None, False),
    (2, (11, 9, 20), "float64", False, "j1", 60, 2e-2, False),
    (2, (11, 9, 20), "float64", False, "jtol", 5, None, False),
    (5, (27, 6, 12), "float64", True, "cg4", 3, None, False),
    (4, (23, 6, 12), "float64", False, "j1", 4, None, True),
    (1, (9, 7, 12), "float64", True, "cg4", 3, None, False),
    # uneven slabs whose launch shapes differ (49 / 48 planes): the workspace layout must still be rank-independent
    (2, (97, 6, 10), "float64", False, "j1", 3, None, False),
    (3, (100, 5, 8), "float64", True, "cg4", 2, None, False),
]


def _reference(shape, dtype, iso, solver, nit, eps, hetero):
    vol = phantom(shape, 3, np.dtype(dtype))
    kw = dict(mu=0.1, omega=1.0, ell=1.0)
    if hetero:
        kw.update(mu=coeff_field(shape, 5, 0.05, 0.1, 0.1).astype(vol.dtype), omega=coeff_field(shape, 6, 0.5, 2.0).astype(vol.dtype))
    hist = []
    ref = oracle.split_bregman_tvd(vol, dim=len(shape), max_num_iter=nit, isotropic=iso, solver=SOLVERS[solver](oracle), eps=eps,
                                   history=hist, **kw)
    return vol, kw, ref, hist


@pytest.mark.parametrize("world,shape,dtype,iso,solver,nit,eps,hetero", CASES)
def test_lock_step_equals_oracle(world, shape, dtype, iso, solver, nit, eps, hetero):
    import darsia_b200 as da
    from darsia_b200.slab_p2p import SlabGroup

    vol, kw, ref, hist = _reference(shape, dtype, iso, solver, nit, eps, hetero)
    with emulated():
        grp = SlabGroup.local(shape, len(shape), vol.dtype, kw["mu"], kw["omega"], kw["ell"], iso, SOLVERS[solver](da), eps=eps,
                              world=world)
        owned = [np.ascontiguousarray(vol[slice(*da.slab_range(shape[0], r, world))]) for r in range(world)]
        grp.setup(owned)
        grp.solve(nit)
        grp.check()
        out = np.concatenate(grp.results(), axis=0)
        h = [grp.history(i) for i in range(world)]
        # a second solve on the same arenas (sequence numbers continue, state reset on the device)
        grp.reset()
        grp.solve(nit)
        grp.check()
        out2 = np.concatenate(grp.results(), axis=0)
    tol = 1e-12 if vol.dtype == np.float64 else 1e-5
    assert rel_l2(out, ref) <= tol
    assert np.array_equal(out, out2)
    assert len(h[0]) == len(hist)
    np.testing.assert_allclose(h[0], hist, rtol=1e-6 if vol.dtype == np.float64 else 1e-3)
    for hi in h[1:]:
        assert hi == h[0]  # every rank derived bitwise the same scalars


@pytest.mark.parametrize("world,shape,dtype,iso,solver,nit,eps,hetero", [CASES[0], CASES[2], CASES[3], CASES[7], CASES[9]])
def test_one_thread_per_slab_equals_oracle(world, shape, dtype, iso, solver, nit, eps, hetero):
    """Every slab driven by its own thread: the flag waits are real."""
    import ctypes as C

    import darsia_b200 as da
    from darsia_b200 import _cabi
    from darsia_b200.slab_p2p import SlabGroup

    vol, kw, ref, hist = _reference(shape, dtype, iso, solver, nit, eps, hetero)
    with emulated() as lib:
        grp = SlabGroup.local(shape, len(shape), vol.dtype, kw["mu"], kw["omega"], kw["ell"], iso, SOLVERS[solver](da), eps=eps,
                              world=world)
        owned = [np.ascontiguousarray(vol[slice(*da.slab_range(shape[0], r, world))]) for r in range(world)]
        grp.setup(owned)  # (lock-step exchange of the static halos)
        # from here on: one private sequence counter and one driver thread per slab
        seqs = [C.c_uint64(grp.seq.value) for _ in range(world)]
        errs = [None] * world

        def drive(i):
            try:
                rk = grp.ranks[i]
                rk.comm.seq = C.pointer(seqs[i])
                rk.problem.max_num_iter = nit
                P = (C.POINTER(_cabi.Problem) * 1)(C.pointer(rk.problem))
                Cm = (C.POINTER(_cabi.SlabComm) * 1)(C.pointer(rk.comm))
                rc = lib.solve_slab(1, P, Cm, None, None, None)
                lib.check(rc, "tvdk_solve_slab")
            except Exception as exc:  # noqa: BLE001
                errs[i] = exc

        threads = [threading.Thread(target=drive, args=(i,)) for i in range(world)]
        for th in threads:
            th.start()
        for th in threads:
            th.join(timeout=300)
        assert all(e is None for e in errs), errs
        assert len({s.value for s in seqs}) == 1  # all ranks walked through the same number of steps
        grp.check()
        out = np.concatenate(grp.results(), axis=0)
        h0 = grp.history(0)
    assert rel_l2(out, ref) <= 1e-12
    assert len(h0) == len(hist)


def test_argument_checks():
    import ctypes as C

    import darsia_b200 as da
    from darsia_b200 import _cabi
    from darsia_b200.slab_p2p import SlabGroup

    with emulated() as lib:
        with pytest.raises(ValueError):
            SlabGroup.local((3, 8, 8), 3, np.float64, 0.1, 1.0, 1.0, False, da.Jacobi(), world=2)  # < 2 planes per slab
        grp = SlabGroup.local((8, 6, 8), 3, np.float64, 0.1, 1.0, 1.0, False, da.Jacobi(), world=2)
        grp.setup([np.zeros((4, 6, 8)), np.zeros((4, 6, 8))])
        rk = grp.ranks[0]
        rk.problem.max_num_iter = 1
        P = (C.POINTER(_cabi.Problem) * 1)(C.pointer(rk.problem))
        Cm = (C.POINTER(_cabi.SlabComm) * 1)(C.pointer(rk.comm))
        saved = rk.problem.geom.z_end
        rk.problem.geom.z_end = saved - 1
        assert lib.solve_slab(1, P, Cm, None, None, None) == -1  # TVDK_ERR_INVALID
        rk.problem.geom.z_end = saved
        rk.problem.f = 1  # outside the arena
        assert lib.solve_slab(1, P, Cm, None, None, None) != 0
        assert b"arena" in lib.last_error()
        # close(): arenas released (peer mappings unmapped first in a distributed group), idempotent, no solve afterwards
        with SlabGroup.local((8, 6, 8), 3, np.float64, 0.1, 1.0, 1.0, False, da.Jacobi(), world=2) as g2:
            g2.setup([np.ones((4, 6, 8)), np.ones((4, 6, 8))])
            g2.solve(2)
            g2.check()
            np.testing.assert_allclose(np.concatenate(g2.results(), axis=0), 1.0, rtol=0, atol=1e-15)  # a constant is a fixed point
        assert all(rk.arena is None for rk in g2.ranks)
        g2.close()
        with pytest.raises(AssertionError):
            g2.solve(1)
