"""GPU, >= 2 devices (NCCL): slab-decomposed solve == single-GPU solve; frame sharding.  Skipped on
a one-GPU box (the same host logic runs under gloo in tests/test_distributed_cpu.py)."""

import os
import socket
import sys
import tempfile

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _ngpu():
    import torch

    return torch.cuda.device_count()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, case, outdir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    import torch
    import torch.distributed as dist

    import bench
    import darsia_b200 as da

    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    try:
        vol = bench.phantom_np(case["shape"], 4, np.dtype(case["dtype"]))
        s, e = da.slab_range(vol.shape[0], rank, world)
        solver = da.CG(maxiter=5, tol=0.0) if case["solver"] == "cg" else da.Jacobi(maxiter=1)
        out, hist = da.split_bregman_tvd_slab(np.ascontiguousarray(vol[s:e]), vol.shape, mu=0.1, omega=1.0, ell=1.0,
                                              dim=3, max_num_iter=case["nit"], isotropic=case["iso"], solver=solver,
                                              return_info=True)
        np.savez(os.path.join(outdir, f"r{rank}.npz"), out=out, hist=np.array(hist))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(_ngpu() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("case", [
    dict(shape=(96, 80, 120), dtype="float64", iso=False, solver="j1", nit=12),
    dict(shape=(96, 80, 120), dtype="float64", iso=True, solver="cg", nit=6),
    dict(shape=(96, 80, 120), dtype="float32", iso=False, solver="j1", nit=12),
])
def test_slab_nccl_equals_single_gpu(case):
    import torch.multiprocessing as mp

    import bench
    import darsia_b200 as da
    from conftest import rel_l2

    world = min(_ngpu(), 4)
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_worker, args=(world, _free_port(), case, d), nprocs=world, join=True)
        parts = [dict(np.load(os.path.join(d, f"r{r}.npz"))) for r in range(world)]
    vol = bench.phantom_np(case["shape"], 4, np.dtype(case["dtype"]))
    solver = da.CG(maxiter=5, tol=0.0) if case["solver"] == "cg" else da.Jacobi(maxiter=1)
    ref = da.split_bregman_tvd(vol, mu=0.1, omega=1.0, ell=1.0, dim=3, max_num_iter=case["nit"],
                               isotropic=case["iso"], solver=solver)
    out = np.concatenate([p["out"] for p in parts], axis=0)
    assert rel_l2(out, ref) <= (1e-12 if vol.dtype == np.float64 else 1e-5)


# ---- the C slab driver (tvdk_solve_slab): halo pushes + ordered sums over peer memory, no NCCL on the data path ----
P2P_CASES = [
    dict(shape=(96, 80, 120), dtype="float64", iso=False, solver="j1", nit=12),
    dict(shape=(96, 80, 120), dtype="float64", iso=True, solver="cg", nit=6),
    dict(shape=(97, 80, 120), dtype="float32", iso=False, solver="j1", nit=12),
    dict(shape=(96, 80, 120), dtype="float64", iso=False, solver="j1", nit=80, eps=5e-3),
]


def _p2p_worker(rank, world, port, case, outdir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    import torch
    import torch.distributed as dist

    import bench
    import darsia_b200 as da
    from darsia_b200.slab_p2p import split_bregman_tvd_slab_p2p

    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    try:
        vol = bench.phantom_np(case["shape"], 4, np.dtype(case["dtype"]))
        s, e = da.slab_range(vol.shape[0], rank, world)
        solver = da.CG(maxiter=5, tol=0.0) if case["solver"] == "cg" else da.Jacobi(maxiter=1)
        out, hist = split_bregman_tvd_slab_p2p(np.ascontiguousarray(vol[s:e]), vol.shape, mu=0.1, omega=1.0, ell=1.0, dim=3,
                                               max_num_iter=case["nit"], isotropic=case["iso"], solver=solver,
                                               eps=case.get("eps"), return_info=True)
        np.savez(os.path.join(outdir, f"r{rank}.npz"), out=out, hist=np.array(hist))
    finally:
        dist.destroy_process_group()


def _single(case):
    import bench
    import darsia_b200 as da

    vol = bench.phantom_np(case["shape"], 4, np.dtype(case["dtype"]))
    solver = da.CG(maxiter=5, tol=0.0) if case["solver"] == "cg" else da.Jacobi(maxiter=1)
    ref, info = da.split_bregman_tvd(vol, mu=0.1, omega=1.0, ell=1.0, dim=3, max_num_iter=case["nit"], isotropic=case["iso"],
                                     solver=solver, eps=case.get("eps"), return_info=True)
    return vol, ref, info


@pytest.mark.skipif(_ngpu() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("case", P2P_CASES)
def test_slab_p2p_one_process_per_gpu_equals_single_gpu(case):
    """One process per GPU: arenas mapped through CUDA IPC, ranks ordered by the mailbox flags."""
    import torch.multiprocessing as mp

    from conftest import rel_l2

    world = min(_ngpu(), 8)
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_p2p_worker, args=(world, _free_port(), case, d), nprocs=world, join=True)
        parts = [dict(np.load(os.path.join(d, f"r{r}.npz"))) for r in range(world)]
    vol, ref, info = _single(case)
    out = np.concatenate([p["out"] for p in parts], axis=0)
    assert rel_l2(out, ref) <= (1e-12 if vol.dtype == np.float64 else 2e-6)
    assert len(parts[0]["hist"]) == int(info.iters[0])
    for p in parts[1:]:
        assert np.array_equal(p["hist"], parts[0]["hist"])  # every rank derived bitwise the same scalars


@pytest.mark.skipif(_ngpu() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("case", P2P_CASES[:3])
def test_slab_p2p_one_process_all_gpus_equals_single_gpu(case):
    """One process drives every GPU in lock step: peer access + event edges between the streams."""
    import darsia_b200 as da
    from conftest import rel_l2
    from darsia_b200.slab_p2p import SlabGroup

    world = min(_ngpu(), 8)
    vol, ref, _ = _single(case)
    solver = da.CG(maxiter=5, tol=0.0) if case["solver"] == "cg" else da.Jacobi(maxiter=1)
    grp = SlabGroup.local(vol.shape, 3, vol.dtype, 0.1, 1.0, 1.0, case["iso"], solver, world=world, devices=list(range(world)))
    grp.setup([np.ascontiguousarray(vol[slice(*da.slab_range(vol.shape[0], r, world))]) for r in range(world)])
    grp.solve(case["nit"])
    grp.check()
    out = np.concatenate(grp.results(), axis=0)
    assert rel_l2(out, ref) <= (1e-12 if vol.dtype == np.float64 else 2e-6)


@pytest.mark.skipif(_ngpu() < 2, reason="needs >= 2 GPUs")
def test_two_devices_from_one_host_thread():
    """per-device launch state.  The isotropic 3-D fused kernel needs the > 48 KB shared-memory opt-in (a
    per-DEVICE function attribute) and ``eps`` uses the polling events (created on one device): both must work when
    the same host thread solves on cuda:0 and then on cuda:1."""
    import torch

    import bench
    import darsia_b200 as da
    from conftest import rel_l2

    vol = bench.phantom_np((40, 48, 96), 7, np.float64)
    kw = dict(mu=0.1, omega=1.0, ell=1.0, dim=3, isotropic=True, max_num_iter=40, eps=2e-2, solver=da.Jacobi(maxiter=1))
    a, ia = da.split_bregman_tvd(vol, device=0, return_info=True, **kw)
    b, ib = da.split_bregman_tvd(vol, device=1, return_info=True, **kw)
    c, ic = da.split_bregman_tvd(vol, device=torch.device("cuda", 0), return_info=True, **kw)
    assert np.array_equal(a, b) and np.array_equal(a, c)
    assert int(ia.iters[0]) == int(ib.iters[0]) == int(ic.iters[0]) <= 40  # (the eps polling ran on both devices)
    vol2 = bench.phantom_np((64, 48, 96), 8, np.float64)  # another shape (another launch configuration) on device 1 first
    kw2 = dict(mu=0.1, omega=1.0, ell=1.0, dim=3, isotropic=False, max_num_iter=5, solver=da.CG(maxiter=3, tol=0.0))
    d1 = da.split_bregman_tvd(vol2, device=1, **kw2)
    d0 = da.split_bregman_tvd(vol2, device=0, **kw2)
    assert rel_l2(d1, d0) == 0.0
