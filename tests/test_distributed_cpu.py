"""CPU, world_size 2 and 3 over gloo: the multi-rank host logic (slab decomposition with halo
exchange + dot-product all-reduce; frame sharding without any collective) with the kernels
running in the emulator.  The decomposed result must equal the single-rank / oracle result."""

import os
import socket
import sys
import tempfile

import numpy as np
import pytest
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, case, outdir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    import torch.distributed as dist

    from emu.harness import emulated
    from golden.generate_golden import coeff_field, phantom

    import darsia_b200 as da

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        with emulated():
            if case["kind"] == "slab":
                vol = phantom(case["shape"], 3, np.dtype(case["dtype"]))
                s, e = da.slab_range(vol.shape[0], rank, world)
                kw = dict(mu=0.1, omega=1.0, ell=1.0)
                if case.get("hetero"):
                    mu = coeff_field(case["shape"], 5, 0.05, 0.1, 0.1).astype(vol.dtype)
                    om = coeff_field(case["shape"], 6, 0.5, 2.0).astype(vol.dtype)
                    kw.update(mu=mu[s:e], omega=om[s:e])
                solver = {"j1": lambda: da.Jacobi(maxiter=1), "j3": lambda: da.Jacobi(maxiter=3),
                          "cg4": lambda: da.CG(maxiter=4, tol=0.0), "cgtol": lambda: da.CG(maxiter=30, tol=1e-3),
                          "jtol": lambda: da.Jacobi(maxiter=12, tol=2e-2)}[case["solver"]]()
                out, hist = da.split_bregman_tvd_slab(np.ascontiguousarray(vol[s:e]), vol.shape, dim=len(case["shape"]),
                                                      max_num_iter=case["nit"], isotropic=case["iso"], solver=solver,
                                                      eps=case.get("eps"), return_info=True, **kw)
                np.savez(os.path.join(outdir, f"r{rank}.npz"), out=out, s=s, e=e, hist=np.array(hist))
            else:
                series = phantom(case["shape"], 4)  # (H, W, T)
                out, (s, e) = da.tvd_series(series, dim=2, mu=0.1, ell=1.0, max_num_iter=5, isotropic=True,
                                            solver=da.CG(maxiter=3, tol=0.0), batch=2, gather=case["gather"])
                np.savez(os.path.join(outdir, f"r{rank}.npz"), out=out, s=s, e=e)
    finally:
        dist.destroy_process_group()


def _run(world, case):
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_worker, args=(world, _free_port(), case, d), nprocs=world, join=True)
        return [dict(np.load(os.path.join(d, f"r{r}.npz"))) for r in range(world)]


SLAB_CASES = [
    (2, dict(kind="slab", shape=(11, 9, 20), dtype="float64", iso=False, solver="j1", nit=6)),
    (3, dict(kind="slab", shape=(13, 9, 20), dtype="float64", iso=True, solver="j1", nit=6, hetero=True)),
    (2, dict(kind="slab", shape=(11, 9, 20), dtype="float64", iso=True, solver="cg4", nit=5)),
    (2, dict(kind="slab", shape=(11, 9, 20), dtype="float64", iso=False, solver="cgtol", nit=4)),
    (2, dict(kind="slab", shape=(11, 9, 20), dtype="float64", iso=False, solver="j3", nit=5)),
    (2, dict(kind="slab", shape=(10, 8, 24), dtype="float32", iso=False, solver="j1", nit=8)),
    (2, dict(kind="slab", shape=(40, 70), dtype="float64", iso=False, solver="j1", nit=6)),
    (2, dict(kind="slab", shape=(11, 9, 20), dtype="float64", iso=False, solver="j1", nit=60, eps=2e-2)),
    (2, dict(kind="slab", shape=(11, 9, 20), dtype="float64", iso=False, solver="jtol", nit=5)),
    # 5 ranks: CG kernels are launched whole (the boundary-first split is for <= 4 ranks), fused Jacobi-1 split
    (5, dict(kind="slab", shape=(27, 6, 12), dtype="float64", iso=True, solver="cg4", nit=3)),
    (5, dict(kind="slab", shape=(46, 6, 12), dtype="float64", iso=False, solver="j1", nit=4)),
]


@pytest.mark.parametrize("world,case", SLAB_CASES)
def test_slab_decomposition_equals_oracle(world, case):
    sys.path.insert(0, HERE)
    import oracle
    from conftest import rel_l2
    from golden.generate_golden import coeff_field, phantom

    parts = _run(world, case)
    vol = phantom(case["shape"], 3, np.dtype(case["dtype"]))
    kw = dict(mu=0.1, omega=1.0, ell=1.0)
    if case.get("hetero"):
        kw.update(mu=coeff_field(case["shape"], 5, 0.05, 0.1, 0.1).astype(vol.dtype),
                  omega=coeff_field(case["shape"], 6, 0.5, 2.0).astype(vol.dtype))
    solver = {"j1": lambda: oracle.Jacobi(maxiter=1), "j3": lambda: oracle.Jacobi(maxiter=3),
              "cg4": lambda: oracle.CG(maxiter=4, tol=0.0), "cgtol": lambda: oracle.CG(maxiter=30, tol=1e-3),
              "jtol": lambda: oracle.Jacobi(maxiter=12, tol=2e-2)}[case["solver"]]()
    hist = []
    ref = oracle.split_bregman_tvd(vol, dim=len(case["shape"]), max_num_iter=case["nit"], isotropic=case["iso"],
                                   solver=solver, eps=case.get("eps"), history=hist, **kw)
    out = np.concatenate([p["out"] for p in parts], axis=0)
    assert [int(p["s"]) for p in parts] == sorted(int(p["s"]) for p in parts)
    tol = 1e-12 if vol.dtype == np.float64 else 1e-5
    assert rel_l2(out, ref) <= tol
    assert len(parts[0]["hist"]) == len(hist)  # same number of outer iterations (eps fired at the same one)
    np.testing.assert_allclose(parts[0]["hist"], hist, rtol=1e-6 if vol.dtype == np.float64 else 1e-3)


@pytest.mark.parametrize("gather", [False, True])
def test_frame_sharding_equals_oracle(gather):
    sys.path.insert(0, HERE)
    import oracle
    from conftest import rel_l2
    from golden.generate_golden import phantom

    case = dict(kind="series", shape=(18, 25, 5), gather=gather)
    parts = _run(2, case)
    series = phantom(case["shape"], 4)
    ref = oracle.split_bregman_tvd(series, mu=0.1, ell=1.0, dim=2, max_num_iter=5, isotropic=True,
                                   solver=oracle.CG(maxiter=3, tol=0.0))
    if gather:
        for p in parts:
            assert rel_l2(p["out"], ref) <= 1e-12
    else:
        out = np.concatenate([p["out"] for p in parts], axis=2)
        assert [(int(p["s"]), int(p["e"])) for p in parts] == [(0, 3), (3, 5)]
        assert rel_l2(out, ref) <= 1e-12
