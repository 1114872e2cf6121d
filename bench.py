#!/usr/bin/env python
"""bench.py -- headline benchmark of the split-Bregman TVD hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

Metric (BASELINE.json): TVD Mvoxel*iter/s -- voxels (all channels/frames) x OUTER Bregman
iterations per second.  One "step" is one whole solve (``tvd()`` with a fixed number of outer
iterations) of one synthetic image batch.

Workload at N = 1 is BASELINE.json configs[1]: a 2-D 4000x7000 3-channel fp64 image, isotropic
TVD, CG(maxiter=10, tol=0) inner solver, 100 outer iterations.  At N > 1 every rank solves its
own image of that shape (frames of a series sharded over ranks: no data-path collective, weak
scaling), the value is the whole-job aggregate over the max-over-ranks time.

Printed JSON (one line, rank 0): see the driver contract; additionally
  roofline     dominant kernel (the CG update kernel), achieved algorithmic GB/s vs measured peak
  cpu_baseline the NumPy oracle (a port of the reference's algorithm) on the host cores
  e2e          same metric through darsia_b200.tvd() with HOST buffers, copies inside the timing
"""

from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: spatial shape, channels, dtype, isotropic, solver (kind, maxiter, tol), outer iterations, hetero
    "cfg2_2d_iso_cg10": dict(shape=(4000, 7000), nb=3, dtype="float64", iso=True, solver=("cg", 10, 0.0),
                             outer=100, hetero=False),
    "cfg1_2d_aniso_j1": dict(shape=(1788, 3180), nb=1, dtype="float64", iso=False, solver=("jacobi", 1, None),
                             outer=100, hetero=False),
    # examples/regularization.py:18-30 as written: Jacobi(maxiter=20), 30 outer iterations, omega = 0.1
    "cfg1_2d_aniso_j20": dict(shape=(1788, 3180), nb=1, dtype="float64", iso=False, solver=("jacobi", 20, None),
                              outer=30, hetero=False, omega=0.1),
    "cfg3_2d_hetero_j1": dict(shape=(4000, 7000), nb=8, dtype="float64", iso=False, solver=("jacobi", 1, None),
                              outer=100, hetero=True),
    "cfg5_3d_aniso_j1_f32": dict(shape=(512, 512, 512), nb=1, dtype="float32", iso=False,
                                 solver=("jacobi", 1, None), outer=50, hetero=False),
    # one volume cut into slabs along axis 0 over the ranks (halo exchange + all-reduce): strong scaling
    "cfg4_3d_slab_cg10": dict(shape=(1024, 1024, 1024), nb=1, dtype="float64", iso=True, solver=("cg", 10, 0.0),
                              outer=20, hetero=False, slab=True),
    "cfg5_3d_slab_j1_f32": dict(shape=(2048, 2048, 512), nb=1, dtype="float32", iso=False,
                                solver=("jacobi", 1, None), outer=50, hetero=False, slab=True),
    "cfg5_3d_slab_j1_f64": dict(shape=(2048, 2048, 512), nb=1, dtype="float64", iso=False,
                                solver=("jacobi", 1, None), outer=50, hetero=False, slab=True),
    # BASELINE configs[2] as specified: 512 frames, per-pixel mu / omega, sharded by frames over the ranks (strong scaling);
    # device-resident value on sub-batches of 8 frames, e2e through the streaming pipeline (tvd_series_stream)
    "cfg3_series512": dict(shape=(4000, 7000), nb=8, dtype="float64", iso=False, solver=("jacobi", 1, None), outer=100,
                           hetero=True, series=512),
    "small": dict(shape=(512, 768), nb=3, dtype="float64", iso=True, solver=("cg", 10, 0.0), outer=10, hetero=False),
}
DEFAULT_WORKLOAD = "cfg2_2d_iso_cg10"
MU, OMEGA, ELL = 0.1, 1.0, 1.0


def algo_words_per_voxel_iter(dim: int, solver, hetero_arrays: int = 0, single_pass: bool = True, big: bool = True) -> float:
    """Algorithmic words per voxel and outer iteration of the IMPLEMENTED formulation
    (compressed state t, q = A p recomputed)."""
    kind, m, _ = solver
    D = dim
    if kind == "jacobi" and m == 1:
        return 3 + 2 * D + hetero_arrays  # read u f t(D), write u t(D)
    if kind == "jacobi":
        # rhs+sweep1: read f t(D) u, write rhs x ; remaining sweeps: 3 words per PASS, a pass doing 4 (2-D) / 2
        # (3-D) sweeps where the count allows (scalar coefficients; per-voxel maps: one sweep per pass) ;
        # shrink: read u+ u t(D), write u t(D)
        rem, passes = m - 1, 0
        while rem > 0:
            s_ = 1 if (hetero_arrays or not single_pass or D != 2 or rem < 2) else (4 if (rem >= 4 and big) else 2)
            rem -= s_
            passes += 1
        return (2 + D + 2) + 3 * passes + (2 + D + 1 + D)
    # CG with a fixed count (tol == 0), single-pass iterations: init (read f t(D) u, write x r) + first pass
    # (read r) + (m-1) fused iterations (read p r x, write p r x; the first of them reads p = r once) + shrink
    # (read x p u t(D), write u t(D))
    if solver[2] == 0.0 and single_pass:
        return (2 + D + 2) + 1 + max(0, 6 * (m - 1) - 1) + (4 + 2 * D)
    # CG with a tolerance: init + per iteration pq 3 (read p r, write p) + update 5 (read p x r, write x r) + shrink
    return (2 + D + 2) + 8 * m + (2 + D + 1 + D)


def phantom_np(shape, seed, dtype):
    """Seeded synthetic image: piecewise-constant blocks + noise, clipped to [0, 1]."""
    rng = np.random.default_rng(seed)
    spatial = shape
    idx = np.indices(spatial, sparse=True)
    cells = [max(2, n // 8) for n in spatial]
    checker = sum((idx[a] // cells[a]) for a in range(len(spatial))) % 3
    base = np.choose(checker, [0.2, 0.5, 0.8])
    img = base + 0.1 * rng.standard_normal(spatial, dtype=np.float32)
    return np.clip(img, 0.0, 1.0).astype(dtype)


def make_host_image(w, seed):
    chans = [phantom_np(w["shape"], seed + c, w["dtype"]) for c in range(w["nb"])]
    return np.stack(chans, axis=-1) if w["nb"] > 1 else chans[0]


def hetero_maps(w, seed):
    rng = np.random.default_rng(seed + 1000)
    shape = w["shape"]
    mu = (0.1 * (0.5 + 0.5 * rng.random(shape, dtype=np.float32))).astype(w["dtype"])
    mu[rng.random(shape, dtype=np.float32) < 0.1] = 0.0  # 10 % zero mask (workflow-style weight*mask)
    om = (0.5 + 1.5 * rng.random(shape, dtype=np.float32)).astype(w["dtype"])
    return mu, om


# --------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi while the timed region runs)
# --------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int) -> None:
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx = float(parts[1])
            except ValueError:
                continue
            for n, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# --------------------------------------------------------------------------------------------
# CPU arm: the UNMODIFIED reference (oracle/_ref, staged by __graft_entry__.build()) on the host cores;
# the NumPy oracle (a port of the same algorithm) only where the reference itself cannot run the workload
# --------------------------------------------------------------------------------------------
def config_of(w, wname):
    """The workload description both arms print (identical dicts => the driver can compare like with like)."""
    return {"workload": wname, "shape": list(w["shape"]), "channels_per_rank": w["nb"], "isotropic": w["iso"],
            "solver": list(w["solver"]), "outer_iterations": w["outer"], "hetero": w["hetero"],
            "l2_policy": "inputs larger than L2 (126 MB)",
            "sharding": ("slabs along axis 0, halo + partial-sum pushes over peer memory" if w.get("slab")
                         else "one image per rank, no collective")}


def cpu_sample(w, budget_voxels=7_000_000):
    """A crop of one channel of the workload such that ~3 outer iterations take 10-30 s."""
    shape = list(w["shape"])
    while np.prod(shape) > budget_voxels:
        a = int(np.argmax(shape))
        shape[a] = max(8, shape[a] // 2)
    return tuple(shape)


def reference_can_run(w):
    """The verbatim reference handles this workload (its isotropic shrinkage is 2-D fp64 with scalar mu)."""
    from oracle import ref_loader

    if not ref_loader.available():
        return False
    if w["iso"] and (len(w["shape"]) != 2 or w["dtype"] != "float64" or w["hetero"]):
        return False
    return True


_CPU_API = {}


def cpu_api(w):
    """(module with the reference's names, kind): the verbatim reference where it can run, else the port."""
    key = reference_can_run(w)
    if key not in _CPU_API:
        if key:
            from oracle import ref_loader

            _CPU_API[key] = (ref_loader.load_reference(), "reference")
        else:
            import oracle

            _CPU_API[key] = (oracle, "port")
    return _CPU_API[key]


def cpu_solve(w, img, n_iter, seed):
    """One call of the reference's entry point on one channel/frame: tvd(img, 'heterogeneous bregman', ...)
    (restoration/tvd.py:132-163; ell = 1.0 is what that entry point passes)."""
    api, _ = cpu_api(w)
    kind, m, tol = w["solver"]
    solver = api.Jacobi(maxiter=m, tol=tol) if kind == "jacobi" else api.CG(maxiter=m, tol=tol)
    kw = dict(weight=MU, omega=w.get("omega", OMEGA), dim=len(img.shape), isotropic=w["iso"], max_num_iter=n_iter, eps=None,
              solver=solver)
    if w["hetero"]:
        mu, om = hetero_maps(dict(w, shape=img.shape), seed)
        kw.update(weight=mu, omega=om)
    t0 = time.perf_counter()
    api.tvd(img, method="heterogeneous bregman", **kw)
    return time.perf_counter() - t0


def cpu_time_marginal(w, seed, n1=1, n2=3):
    """Seconds per outer iteration on the bounded sample: marginal cost (t(n2) - t(n1)) / (n2 - n1) of two calls
    (absorbs the per-call Numba compile of the reference; falls back to t(n2)/n2 if noise makes it <= 0)."""
    shape = cpu_sample(w)
    img = phantom_np(shape, seed, w["dtype"])
    cpu_solve(w, phantom_np(tuple(16 for _ in shape), 0, w["dtype"]), 1, 0)  # imports, BLAS pool
    ts = [cpu_solve(w, img, n, seed) for n in (n1, n2)]
    per_iter = (ts[1] - ts[0]) / (n2 - n1)
    if per_iter <= 0:
        per_iter = ts[1] / n2
    return shape, per_iter, ts


def blas_threads():
    try:
        from threadpoolctl import threadpool_info

        return max([p.get("num_threads", 1) for p in threadpool_info()] or [1])
    except Exception:
        return None


def _ref_worker(job):
    """One channel/frame of the sample through the reference; returns (voxels*iters, seconds)."""
    w, shape, seed, n_iter = job
    img = phantom_np(shape, seed, w["dtype"])
    dt = cpu_solve(w, img, n_iter, seed)
    return int(np.prod(shape)) * n_iter, dt


def run_reference_arm(args, w, wname):
    """--impl reference: the reference's own CPU implementation of the path on the host cores -- the UNMODIFIED
    DarSIA files (oracle/_ref, executed verbatim by oracle/ref_loader.py, Numba shrinkage included) through
    darsia.tvd(), on the GPU arm's workload: full-size channels, same solver, same config dict.  The reference is
    single-threaded per image (element-wise NumPy, only np.dot is threaded); to use the host's cores the way a
    user could, the channels/frames of the workload run in parallel processes.  One step = ONE outer iteration
    per channel (a bounded sample of the 100 the GPU arm runs; the cost per outer iteration is constant)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp

    _, kind = cpu_api(w)
    nproc = max(1, min(w["nb"] if not w.get("slab") else 1, (os.cpu_count() or 2) // 2, 8))
    n_iter = 1
    full = tuple(w["shape"])
    if w.get("slab"):  # a 1024^3 volume does not fit the host with NumPy temporaries: 256^3-class sub-volume
        full = cpu_sample(w, 20_000_000)
    budget_s = float(os.environ.get("BENCH_REF_BUDGET_S", "240"))
    ctx = mp.get_context("fork")
    vals, secs = [], []
    tiny = tuple(max(16, n // 16) for n in full)
    with ctx.Pool(nproc) as pool:
        pool.map(_ref_worker, [(w, tiny, 0, 1)] * nproc)  # start-up costs (imports, Numba/BLAS initialisation)
        # the first warm-up step runs the FULL sample and sizes the timed steps: if K of them would exceed the budget the
        # sample becomes a band of full-width rows (cost per voxel is size-independent for this streaming code)
        shape = full
        for s_ in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            warm = s_ < args.warmup
            shp = shape if (not warm or s_ == 0) else tiny
            res = pool.map(_ref_worker, [(w, shp, 2 + 10 * s_ + c, n_iter) for c in range(nproc)])
            wall = time.perf_counter() - t0
            if s_ == 0 and wall * args.steps > budget_s and (warm or args.warmup == 0):
                rows = max(64, int(full[0] * budget_s / (wall * args.steps)) // 8 * 8)
                shape = (min(full[0], rows),) + full[1:]
            if not warm:
                vals.append(sum(r[0] for r in res) / wall / 1e6)
                secs.append(wall)
    v = float(np.mean(vals))
    what = ("the UNMODIFIED DarSIA reference (oracle/_ref, verbatim, Numba shrinkage incl. its per-call JIT) through darsia.tvd()"
            if kind == "reference" else
            "NumPy oracle (port of the reference algorithm: the reference itself cannot run this workload, SURVEY Q3)")
    sample = (f"{what}: {nproc} channel(s)/frame(s) of shape {tuple(shape)} (workload shape {tuple(w['shape'])}) in {nproc} parallel "
              f"processes, {n_iter} outer iteration per channel and step; element-wise NumPy is single-threaded per process, "
              f"BLAS threads={blas_threads()}, host cpus={os.cpu_count()}")
    line = {
        "impl": "reference", "metric": "tvd_mvoxel_iter_per_s", "value": v, "unit": "Mvoxel*iter/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * float(np.mean(secs)), "higher_is_better": True,
        "scaling": "strong" if w.get("slab") else "weak", "vs_baseline": None,
        "dtype": "f64" if w["dtype"] == "float64" else "f32", "data": "synthetic",
        "config": config_of(w, wname),
        "cpu_baseline": {"value": v, "unit": "Mvoxel*iter/s", "cores": nproc, "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": "Mvoxel*iter/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------
def ncu_traffic(kernel_key):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture (profiles/), if any."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        with open(path) as fh:
            return json.load(fh).get(kernel_key)
    except Exception:
        return None


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def run_ours(args, w, wname):
    import torch
    import torch.distributed as dist

    import darsia_b200 as da
    from darsia_b200 import _cabi
    from darsia_b200.engine import Engine, SolverSpec

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local])
        torch.cuda.synchronize(dev)

    lib = _cabi.get_library()
    dim = len(w["shape"])
    kind, m, tol = w["solver"]
    spec = SolverSpec(kind=_cabi.TVDK_SOLVER_CG if kind == "cg" else _cabi.TVDK_SOLVER_JACOBI, maxiter=m, tol=tol)
    np_dtype = np.dtype(w["dtype"])
    nvox_step = int(np.prod(w["shape"])) * w["nb"]  # per rank
    outer = w["outer"]

    # ---- synthetic inputs: one host image per rank (seeded), resident copy on the device ----
    host_img = make_host_image(w, seed=2 + 100 * rank)
    pinned_in = torch.from_numpy(host_img).pin_memory()
    pinned_np = pinned_in.numpy()
    eng = Engine(w["shape"], nb=w["nb"], dtype=np_dtype, device=dev)
    mu, om = MU, w.get("omega", OMEGA)
    n_het = 0
    if w["hetero"]:
        mu_h, om_h = hetero_maps(w, seed=2)
        mu, _k1 = eng.coef(mu_h)[1], None
        om = eng.coef(om_h)[1]
        n_het = 2
        mu_host, om_host = mu_h, om_h
    f = eng.upload(pinned_np)
    u = eng.empty_fields(1)
    t = eng.empty_fields(dim)

    def device_step():
        u.copy_(f)  # d2d reset of the iterate (plumbing, outside the algorithmic bytes; < 1 % of a step)
        t.zero_()
        eng.solve(f, u, t, mu, om, ELL, w["iso"], spec, outer, None, sync=False)

    # ---- warm-up, then K timed steps with CUDA events on the launching stream ----
    for _ in range(args.warmup):
        device_step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    lib.launch_count(1)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        device_step()
    ev1.record()
    barrier()
    launches = int(lib.launch_count(1))
    clocks = sampler.stop() if rank == 0 else None
    ms = ev0.elapsed_time(ev1)
    tm = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    ms_total = float(tm.item())
    ms_per_step = ms_total / args.steps
    value = world * nvox_step * outer / (ms_per_step * 1e-3) / 1e6

    # ---- parity spot-check of what was just timed (cheap, not timed): finite result, the single-pass CG diagnostic,
    #      and the same solver at the config's iteration counts on a reduced shape against the oracle ----
    res = eng.unpack(u)
    assert bool(torch.isfinite(res).all()), "non-finite result"
    info_full = eng.solve(f, u.copy_(f), t.zero_(), mu, om, ELL, w["iso"], spec, min(outer, 5), None, sync=True)
    cg_pred_dev = float(np.max(info_full.cg_pred_dev)) if info_full.cg_pred_dev is not None else None
    parity = None
    if rank == 0 and not args.no_cpu and not w["hetero"]:
        import oracle

        red = tuple(max(24, n // 16) + 1 for n in w["shape"])  # odd extents, several tiles per axis
        small = phantom_np(red, 77, w["dtype"])
        mk = (lambda mod: mod.CG(maxiter=m, tol=tol)) if kind == "cg" else (lambda mod: mod.Jacobi(maxiter=m, tol=tol))
        kws = dict(mu=MU, omega=w.get("omega", OMEGA), ell=ELL, dim=dim, isotropic=w["iso"], max_num_iter=outer)
        got = da.split_bregman_tvd(small, solver=mk(da), device=dev, **kws)
        want = oracle.split_bregman_tvd(small, solver=mk(oracle), **kws)
        parity = {"rel_l2_vs_oracle": float(np.linalg.norm(got - want) / np.linalg.norm(want)), "shape": list(red),
                  "outer_iterations": outer, "bar": 1e-10 if np_dtype == np.float64 else 1e-5}

    # ---- dominant kernel alone: average launch duration with CUDA events ----
    W = np_dtype.itemsize
    g = eng.geom
    ws = eng.op_workspace()
    sums = torch.zeros(w["nb"], dtype=torch.float64, device=dev)
    coef = torch.full((w["nb"],), 1e-3, dtype=torch.float64, device=dev)
    omc, _ = eng.coef(om)
    elc, _ = eng.coef(ELL)
    muc, _ = eng.coef(mu)
    if kind == "cg" and tol == 0.0:
        dom_name, dom_words, dom_key = ("cg_fused (x += a p; r -= a A p; p = r + b p; r.r, p.Ap, r.Ap, Ap.Ap)", 6,
                                        "cg_fused")
        p_, x_, r_ = eng.upload(pinned_np), eng.zero_fields(1), eng.upload(pinned_np)
        p2_, r2_ = eng.empty_fields(1), eng.empty_fields(1)
        sums = torch.zeros(4 * w["nb"], dtype=torch.float64, device=dev)

        def dom():
            return lib.cg_fused(C.byref(g), C.byref(omc), C.byref(elc), p_.data_ptr(), r_.data_ptr(), p2_.data_ptr(),
                                r2_.data_ptr(), x_.data_ptr(), coef.data_ptr(), coef.data_ptr(), sums.data_ptr(),
                                ws.data_ptr(), ws.numel(), eng.stream())
    elif kind == "cg":
        dom_name, dom_words, dom_key = "cg_update (x += a p; r -= a A p; r.r)", 5, "cg_update"
        p_, x_, r_ = eng.upload(pinned_np), eng.zero_fields(1), eng.zero_fields(1)

        def dom():
            return lib.cg_update(C.byref(g), C.byref(omc), C.byref(elc), p_.data_ptr(), x_.data_ptr(), r_.data_ptr(),
                                 coef.data_ptr(), sums.data_ptr(), ws.data_ptr(), ws.numel(), eng.stream())
    elif kind == "jacobi" and m >= 2:
        nsw = 1 if (n_het or dim != 2) else (4 if (m >= 5 and int(np.prod(w["shape"])) >= (1 << 23)) else 2)
        dom_name, dom_words, dom_key = f"jacobi sweeps, {nsw} per pass (x = rhs/diag + N(x)/(diag/ell))", 3, "jacobi_sweeps"
        xa_, xb_ = eng.upload(pinned_np), eng.empty_fields(1)

        def dom():
            return lib.jacobi_sweeps(C.byref(g), C.byref(omc), C.byref(elc), xa_.data_ptr(), f.data_ptr(), xb_.data_ptr(),
                                     nsw, None, ws.data_ptr(), ws.numel(), eng.stream())
    else:
        dom_name, dom_words, dom_key = "fused_jacobi1 (rhs+sweep+shrink+update)", 3 + 2 * dim + n_het, "fused_jacobi1"
        u2, t2 = eng.empty_fields(1), eng.empty_fields(dim)
        t.zero_()

        def dom():
            return lib.fused_jacobi1(C.byref(g), 1 if w["iso"] else 0, C.byref(muc), C.byref(omc), C.byref(elc),
                                     f.data_ptr(), u.data_ptr(), u2.data_ptr(), t.data_ptr(), t2.data_ptr(),
                                     eng.batch_elems, sums.data_ptr(), ws.data_ptr(), ws.numel(), eng.stream())
    for _ in range(3):
        lib.check(dom(), "dominant kernel")
    torch.cuda.synchronize(dev)
    reps = 20
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k0.record()
    for _ in range(reps):
        dom()
    k1.record()
    torch.cuda.synchronize(dev)
    dom_ms = k0.elapsed_time(k1) / reps
    peak, peak_src = measured_peak()
    dom_bytes = dom_words * W * nvox_step
    achieved = dom_bytes / (dom_ms * 1e-3) / 1e9
    # ---- the same kernel family INSIDE one more (untimed) step: every launch bracketed by CUDA events on its stream
    #      (tvdk_profile_arm); sustained clocks and the cache state of the real sequence, cross-checks ncu's share ----
    in_step = None
    pattern = {"cg_fused": b"KCgFused", "fused_jacobi1": b"KFusedJ1", "jacobi_sweeps": b"KSweep"}.get(dom_key)
    if pattern is not None:
        lib.check(lib.profile_arm(pattern), "tvdk_profile_arm")
        device_step()
        n_l, tot_ms = C.c_int64(0), C.c_double(0.0)
        lib.check(lib.profile_read(C.byref(n_l), C.byref(tot_ms)), "tvdk_profile_read")
        if n_l.value:
            avg = tot_ms.value / n_l.value
            in_step = {"kernel_type_contains": pattern.decode(), "launches": int(n_l.value), "avg_launch_ms": avg,
                       "share_of_step": tot_ms.value / ms_per_step,
                       "note": "one extra instrumented step after the timed ones; sustained clocks (see clocks.reasons)"}
            if dom_key == "cg_fused":  # every launch of the family moves dom_words per voxel
                in_step["achieved"] = dom_bytes / (avg * 1e-3) / 1e9
                in_step["frac_of_burst_peak"] = in_step["achieved"] / peak
    words_iter = algo_words_per_voxel_iter(dim, w["solver"], n_het, big=int(np.prod(w["shape"])) >= (1 << 23))
    step_bytes = words_iter * W * nvox_step * outer
    step_gbs = step_bytes / (ms_per_step * 1e-3) / 1e9
    survey_words = {"cg": 8 + 5 * dim + 10 * m, "jacobi": (3 + 4 * dim) if m == 1 else (3 * m + 2 + 5 * dim)}[kind] + n_het
    roofline = {
        "bound": "hbm", "kernel": dom_name, "achieved": achieved, "peak": peak, "unit": "GB/s",
        "frac": achieved / peak, "traffic": ncu_traffic(f"{dom_key}:{wname}"), "peak_source": peak_src,
        "traffic_source": "profiles/ncu_traffic.json: dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture "
                          "of this kernel on this workload (committed; not re-measured in this run)",
        "frac_of_8TBs_spec": achieved / 8000.0,  # north_star quotes the ~8 TB/s nominal HBM3e figure
        "algo_bytes_per_launch": dom_bytes, "avg_launch_ms": dom_ms,
        "timed": "the kernel alone, 20 back-to-back launches after the timed steps (burst, like the measured peak)",
        "in_step": in_step,
        "whole_step": {"algo_words_per_voxel_iter": words_iter, "achieved": step_gbs, "frac": step_gbs / peak,
                       "survey_words_per_voxel_iter": survey_words,
                       "frac_vs_survey_bytes": survey_words * W * nvox_step * outer / (ms_per_step * 1e-3) / 1e9 / peak},
    }

    # ---- e2e: public API, host buffers in, host array out, copies inside the timed region ----
    solver_obj = (da.CG(maxiter=m, tol=tol) if kind == "cg" else da.Jacobi(maxiter=m, tol=tol))
    kw = dict(method="heterogeneous bregman", weight=(mu_host if w["hetero"] else MU),
              omega=(om_host if w["hetero"] else w.get("omega", OMEGA)), regularization=ELL, dim=dim, isotropic=w["iso"],
              max_num_iter=outer, eps=None, solver=solver_obj)
    e2e_steps = max(1, min(args.steps, 3))
    # steady state of a caller that processes image after image: three warm-up calls fill torch's
    # device and pinned-host allocator caches (a fresh 672 MB cudaHostAlloc alone costs ~0.2 s)
    for _ in range(3):
        out = da.tvd(pinned_np, **kw)
        del out
    barrier()
    os.environ["DARSIA_B200_TIMING"] = "1"
    t0 = time.perf_counter()
    out = None
    e2e_each = []
    for _ in range(e2e_steps):
        del out  # a streaming caller consumes one result before asking for the next (the pinned block is reused)
        t1 = time.perf_counter()
        out = da.tvd(pinned_np, **kw)
        e2e_each.append(round(time.perf_counter() - t1, 4))
    torch.cuda.synchronize(dev)
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    os.environ.pop("DARSIA_B200_TIMING", None)
    import importlib

    _sbt = importlib.import_module("darsia_b200.restoration.split_bregman_tvd")
    e2e_breakdown = {k: round(v, 4) for k, v in _sbt.LAST_TIMINGS.items()}
    e2e_breakdown["each_call_s"] = e2e_each
    e2e_breakdown["input_is_pinned"] = bool(torch.from_numpy(pinned_np).is_pinned())
    # the same call on a PAGEABLE ndarray (what a drop-in caller of darsia.tvd passes)
    da.tvd(host_img, **kw)
    torch.cuda.synchronize(dev)
    tp0 = time.perf_counter()
    outp = None
    for _ in range(e2e_steps):
        del outp  # (as above: a result kept alive forces a fresh 672 MB cudaHostAlloc, ~0.2 s, for the next one)
        outp = da.tvd(host_img, **kw)
    torch.cuda.synchronize(dev)
    e2e_pageable_s = (time.perf_counter() - tp0) / e2e_steps
    del outp
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_val = world * nvox_step * outer / float(te.item()) / 1e6
    h2d = host_img.nbytes + (mu_host.nbytes + om_host.nbytes if w["hetero"] else 0)
    d2h = out.nbytes + w["nb"] * C.sizeof(_cabi.Stat)
    # the e2e result must be the device-timed result (same inputs, deterministic kernels)
    same = bool(np.array_equal(out.reshape(-1), res.cpu().numpy().reshape(-1)))

    # ---- CPU baseline on rank 0 (N = 1 only) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        shape_s, per_iter, ts = cpu_time_marginal(w, seed=2)
        _, cpu_kind = cpu_api(w)
        cpu = {
            "value": float(np.prod(shape_s)) / per_iter / 1e6, "unit": "Mvoxel*iter/s", "cores": 1, "kind": cpu_kind,
            "sample": (("the UNMODIFIED DarSIA reference (oracle/_ref, executed verbatim) through darsia.tvd()" if cpu_kind == "reference"
                        else "NumPy oracle (port of the reference algorithm)") +
                       f" on one channel cropped to {tuple(shape_s)}: marginal cost (t(3 its)-t(1 it))/2 = "
                       f"{per_iter:.3f} s per outer iteration; raw {ts[1] / 3:.3f} s/iter; element-wise NumPy is "
                       f"single-threaded, host has {os.cpu_count()} cpus, BLAS threads={blas_threads()}"),
        }
    # ---- the slab-decomposed configs (one volume over all ranks: halo + partial-sum pushes), strong scaling ----
    slab = None
    if wname == DEFAULT_WORKLOAD and not args.no_slab:
        del f, u, t, res, out, pinned_in, dom  # (dom's closure holds the dominant-kernel buffers)
        eng = None
        import gc

        gc.collect()
        torch.cuda.empty_cache()
        slab = {}
        for sname in ("cfg4_3d_slab_cg10", "cfg5_3d_slab_j1_f32"):
            slab[sname] = slab_measure(WORKLOADS[sname], sname, rank, world, local, dev, steps=2, warmup=1,
                                       want_e2e=world > 1, barrier=barrier)
            gc.collect()
            torch.cuda.empty_cache()
    if world > 1:
        dist.destroy_process_group()
    if rank != 0:
        return
    line = {
        "metric": "tvd_mvoxel_iter_per_s", "value": value, "unit": "Mvoxel*iter/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64" if np_dtype == np.float64 else "f32",
        "data": "synthetic (seeded block phantom + noise; random-free weights: scalar mu/omega/ell)",
        "config": config_of(w, wname),
        "roofline": roofline,
        "cpu_baseline": cpu,
        "e2e": {"value": e2e_val, "unit": "Mvoxel*iter/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "s_per_step": float(te.item()), "matches_device_timed_result": same,
                "breakdown_s": e2e_breakdown,
                "pageable_input": {"value": world * nvox_step * outer / e2e_pageable_s / 1e6, "s_per_step": e2e_pageable_s}},
        "parity": parity, "cg_pred_dev": cg_pred_dev,
        "gpu_launches": launches,
        "clocks": clocks,
    }
    if slab is not None:
        line["slab"] = slab
    print(json.dumps(line))


def phantom_torch(shape, seed, tdtype, dev):
    """Device-side synthetic volume of the same kind as phantom_np (blocks + noise, clipped): used where a host
    array of the whole volume would take minutes to generate (1024^3 on one GPU)."""
    import torch

    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    cells = [max(2, n // 8) for n in shape]
    checker = None
    for a, n in enumerate(shape):
        ix = (torch.arange(n, device=dev) // cells[a]).reshape([-1 if i == a else 1 for i in range(len(shape))])
        checker = ix if checker is None else checker + ix
    base = torch.tensor([0.2, 0.5, 0.8], device=dev, dtype=torch.float32)[checker % 3]
    base += 0.1 * torch.randn(shape, generator=g, device=dev, dtype=torch.float32)
    return base.clamp_(0.0, 1.0).to(tdtype)


def slab_measure(w, wname, rank, world, local, dev, steps, warmup, want_e2e, barrier, transport="p2p"):
    """One volume cut into slabs over the ranks (BASELINE configs 4/5), through tvdk_solve_slab: device-resident
    loop timed with CUDA events (max over ranks), the same slab solved WITHOUT neighbours (what the kernels alone
    cost), a parity check of the decomposition against one GPU on a reduced volume, optionally e2e with host buffers."""
    import torch
    import torch.distributed as dist

    import darsia_b200 as da
    from darsia_b200.slab_p2p import SlabGroup

    shape = tuple(w["shape"])
    dim = len(shape)
    kind, m, tol = w["solver"]
    np_dtype = np.dtype(w["dtype"])
    tdtype = torch.float64 if np_dtype == np.float64 else torch.float32
    mk = lambda: (da.CG(maxiter=m, tol=tol) if kind == "cg" else da.Jacobi(maxiter=m, tol=tol))  # noqa: E731
    s, e = da.slab_range(shape[0], rank, world)
    outer = w["outer"]
    nvox = int(np.prod(shape))
    W = np_dtype.itemsize

    def reduce_max(x):
        tx = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tx, op=dist.ReduceOp.MAX)
        return float(tx.item())

    def timed(grp, n):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record()
        for _ in range(n):
            grp.reset()
            grp.solve(outer)
        ev1.record()
        barrier()
        return reduce_max(ev0.elapsed_time(ev1)) / n

    def everyone_has(nbytes, what):
        """True when EVERY rank has `nbytes` of device memory free (the same answer on all ranks, so that a stage is
        skipped by all of them or by none: a rank that ran out of memory alone would leave the others in a barrier)."""
        free = torch.cuda.mem_get_info(dev)[0]
        ok = torch.tensor([1.0 if free > nbytes + (2 << 30) else 0.0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if not bool(ok.item()):
            skipped.append(f"{what}: needs {nbytes / 2**30:.1f} GiB free on every GPU")
        return bool(ok.item())

    skipped = []
    slab_field = (e - s + 3) * int(np.prod(shape[1:])) * W  # one field of this rank's slab incl. ghost planes
    n_fields = 2 + dim + (5 if kind == "cg" else 2)  # f, u, t + the solver's workspace (upper bound)
    grp = SlabGroup.distributed(shape, dim, np_dtype, MU, OMEGA, ELL, w["iso"], mk(), device=dev)
    owned_dev = phantom_torch((e - s,) + shape[1:], 4 + rank, tdtype, dev)
    grp.setup([owned_dev])
    del owned_dev
    for _ in range(warmup):
        grp.reset()
        grp.solve(outer)
    grp.lib.launch_count(1)
    ms_per_step = timed(grp, steps)
    launches = int(grp.lib.launch_count(1)) // steps
    grp.check()
    value = nvox * outer / (ms_per_step * 1e-3) / 1e6
    fin = bool(torch.isfinite(grp.ranks[0].field("u")).all())
    words = algo_words_per_voxel_iter(dim, w["solver"], 0, single_pass=False)  # 3-D / slabs: two-kernel CG iterations
    peak, peak_src = measured_peak()
    gbs = words * W * nvox * outer / (ms_per_step * 1e-3) / 1e9
    rec = {"value": value, "unit": "Mvoxel*iter/s", "ms_per_step": ms_per_step, "n_gpus": world, "scaling": "strong",
           "config": config_of(w, wname), "finite": fin, "steps": steps, "warmup": warmup,
           "roofline_per_gpu": {"achieved": gbs / world, "peak": peak, "unit": "GB/s", "frac": gbs / world / peak,
                                "algo_words_per_voxel_iter": words, "peak_source": peak_src},
           "gpu_launches_per_step": launches,
           "transport": "peer-memory pushes + flags (tvdk_solve_slab); no NCCL call inside the timed region"}
    # comm steps per solve: 2 (handshake, ||f||) + per outer iteration 1 (fused J1) / 2m + 2 (CG) / K + 1 (Jacobi K)
    per_it = 1 if (kind == "jacobi" and m == 1) else (2 * m + 2 if kind == "cg" else m + 1)
    rec["comm_steps_per_solve"] = 2 + per_it * outer
    if world > 1:
        # the same slab without neighbours (kernels + the same tiny launches, no peer traffic, no waiting on anybody)
        grp.close()
        del grp
        torch.cuda.empty_cache()
        alone = SlabGroup.local((e - s,) + shape[1:], dim, np_dtype, MU, OMEGA, ELL, w["iso"], mk(), world=1, devices=[dev])
        alone.setup([phantom_torch((e - s,) + shape[1:], 4 + rank, tdtype, dev)])
        alone.reset()
        alone.solve(outer)
        ms_alone = timed(alone, steps)
        alone.close()
        del alone
        torch.cuda.empty_cache()
        rec["breakdown_ms"] = {"total": ms_per_step, "own_slab_without_neighbours": ms_alone,
                               "halo_pushes_and_flag_waits": ms_per_step - ms_alone,
                               "per_comm_step_us": 1e3 * (ms_per_step - ms_alone) / rec["comm_steps_per_solve"]}
        # decomposition == one GPU on a reduced volume (same per-voxel arithmetic; only the order of the partial sums differs)
        red = (max(4 * world, 64),) + tuple(max(48, n // 8) for n in shape[1:])
        small = phantom_np(red, 11, w["dtype"])
        rs, re_ = da.slab_range(red[0], rank, world)
        g2 = SlabGroup.distributed(red, dim, np_dtype, MU, OMEGA, ELL, w["iso"], mk(), device=dev)
        g2.setup([np.ascontiguousarray(small[rs:re_])])
        g2.solve(outer)
        mine = torch.from_numpy(g2.results()[0]).to(dev)
        g2.check()
        parts = [torch.empty((da.slab_range(red[0], r, world)[1] - da.slab_range(red[0], r, world)[0],) + red[1:], dtype=mine.dtype,
                             device=dev) for r in range(world)]
        dist.all_gather(parts, mine)
        if rank == 0:
            whole = torch.cat(parts, dim=0).cpu().numpy()
            single = da.split_bregman_tvd(small, mu=MU, omega=OMEGA, ell=ELL, dim=dim, max_num_iter=outer, isotropic=w["iso"],
                                          solver=mk(), device=dev)
            rec["parity_vs_one_gpu"] = {"rel_l2": float(np.linalg.norm(whole - single) / np.linalg.norm(single)),
                                        "shape": list(red), "outer_iterations": outer}
        g2.close()
        del g2, parts, mine
        torch.cuda.empty_cache()
        # SURVEY 8d parity protocol (iii): the decomposed result == the single-GPU result AT FULL SIZE.  Every rank repeats
        # its solve, the owned planes are gathered on rank 0, which solves the whole volume alone (it fits: 9 fields of
        # 8.6 GB) from the same seeded slabs and compares on the device.
        equal_slabs = len({da.slab_range(shape[0], r, world)[1] - da.slab_range(shape[0], r, world)[0] for r in range(world)}) == 1
        if equal_slabs and everyone_has((n_fields + 2) * nvox * W if rank == 0 else (n_fields + 1) * slab_field, "full-size parity"):
            gfull = SlabGroup.distributed(shape, dim, np_dtype, MU, OMEGA, ELL, w["iso"], mk(), device=dev)
            gfull.setup([phantom_torch((e - s,) + shape[1:], 4 + rank, tdtype, dev)])
            gfull.solve(outer)
            gfull.check()
            rk = gfull.ranks[0]
            mine_f = rk.field("u")[rk.plane_elems * 1:rk.plane_elems * (1 + (e - s))].clone()
            gfull.close()
            del gfull, rk
            torch.cuda.empty_cache()
            gathered = [torch.empty_like(mine_f) for _ in range(world)] if rank == 0 else None
            dist.gather(mine_f, gathered, dst=0)
            del mine_f
            if rank == 0:
                whole_in = torch.cat([phantom_torch((da.slab_range(shape[0], r, world)[1] - da.slab_range(shape[0], r, world)[0],)
                                                    + shape[1:], 4 + r, tdtype, dev) for r in range(world)], dim=0)
                one = SlabGroup.local(shape, dim, np_dtype, MU, OMEGA, ELL, w["iso"], mk(), world=1, devices=[dev])
                one.setup([whole_in])
                del whole_in
                one.solve(outer)
                r1 = one.ranks[0]
                ref_f = r1.field("u")[r1.plane_elems * 1:r1.plane_elems * (1 + shape[0])]
                num, den = 0.0, 0.0
                o = 0
                for part in gathered:  # slab by slab: no second copy of the volume
                    seg = ref_f[o:o + part.numel()]
                    num += float(torch.sum((part.double() - seg.double()) ** 2))
                    den += float(torch.sum(seg.double() ** 2))
                    o += part.numel()
                rec["parity_full_size_vs_one_gpu"] = {"rel_l2": float(np.sqrt(num / den)), "shape": list(shape),
                                                      "outer_iterations": outer}
                one.close()
                del one, r1, ref_f, seg, part  # (seg is a view of one's arena: it alone would keep 9 fields alive)
            del gathered
            torch.cuda.empty_cache()
            barrier()
        if want_e2e and everyone_has((n_fields + 2) * slab_field, "e2e"):
            # (host planes = the device phantom copied out once: numpy takes ~30 s to synthesise 0.5 G voxels)
            owned_t = torch.empty((e - s,) + shape[1:], dtype=tdtype, pin_memory=True)
            owned_t.copy_(phantom_torch((e - s,) + shape[1:], 4 + rank, tdtype, dev))
            torch.cuda.synchronize(dev)
            owned = owned_t.numpy()
            g3 = SlabGroup.distributed(shape, dim, np_dtype, MU, OMEGA, ELL, w["iso"], mk(), device=dev)
            g3.setup([owned])
            g3.solve(outer)
            g3.results()  # warm-up of the allocators / pinned staging
            barrier()
            t0 = time.perf_counter()
            g3.setup([owned])
            g3.solve(outer)
            out = g3.results(w["dtype"])[0]
            torch.cuda.synchronize(dev)
            te = reduce_max(time.perf_counter() - t0)
            rec["e2e"] = {"value": nvox * outer / te / 1e6, "unit": "Mvoxel*iter/s", "s_per_step": te,
                          "h2d_bytes_per_step": int(owned.nbytes), "d2h_bytes_per_step": int(out.nbytes),
                          "input": "owned planes in pinned host memory"}
            g3.close()
            del g3, owned, owned_t, out
            torch.cuda.empty_cache()
    else:
        grp.close()
    if skipped:
        rec["skipped"] = skipped
    return rec


def run_slab(args, w, wname):
    """One volume, slab-decomposed over the ranks (BASELINE configs 4/5): strong scaling."""
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local])
        torch.cuda.synchronize(dev)

    if args.transport == "nccl":
        return run_slab_nccl(args, w, wname, rank, world, local, dev, barrier)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    rec = slab_measure(w, wname, rank, world, local, dev, args.steps, args.warmup, want_e2e=world > 1, barrier=barrier)
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        dist.destroy_process_group()
    if rank != 0:
        return
    e2e = rec.pop("e2e", None)
    rl = rec.pop("roofline_per_gpu")
    line = {
        "metric": "tvd_mvoxel_iter_per_s", "value": rec.pop("value"), "unit": "Mvoxel*iter/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": rec.pop("ms_per_step"), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64" if w["dtype"] == "float64" else "f32", "data": "synthetic",
        "config": rec.pop("config"),
        "roofline": dict(rl, bound="hbm", kernel="whole step (all kernels + halo pushes + flag waits)", traffic=None, per_gpu=True),
        "cpu_baseline": None, "e2e": e2e, "gpu_launches": rec.pop("gpu_launches_per_step") * args.steps, "clocks": clocks,
        "slab": rec,
    }
    print(json.dumps(line))


def run_slab_nccl(args, w, wname, rank, world, local, dev, barrier):
    """The round-1 slab driver (Python loop, NCCL isend/irecv + all_reduce), kept for A/B runs: --transport nccl."""
    import torch
    import torch.distributed as dist

    import darsia_b200 as da
    from darsia_b200.slab import SlabSolver

    shape = tuple(w["shape"])
    dim = len(shape)
    kind, m, tol = w["solver"]
    solver = da.CG(maxiter=m, tol=tol) if kind == "cg" else da.Jacobi(maxiter=m, tol=tol)
    s, e = da.slab_range(shape[0], rank, world)
    owned = phantom_np((e - s,) + shape[1:], 4 + rank, w["dtype"])
    sol = SlabSolver(shape, dim, np.dtype(w["dtype"]), MU, OMEGA, ELL, w["iso"], solver, device=dev)
    sol.setup(owned)
    outer = w["outer"]
    nvox = int(np.prod(shape))
    for _ in range(args.warmup):
        sol.reset()
        sol.iterate(outer)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        sol.reset()
        sol.iterate(outer)
    ev1.record()
    barrier()
    tm = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        dist.destroy_process_group()
    ms_per_step = float(tm.item()) / args.steps
    if rank != 0:
        return
    W = np.dtype(w["dtype"]).itemsize
    words = algo_words_per_voxel_iter(dim, w["solver"], 0, single_pass=False)
    peak, peak_src = measured_peak()
    gbs = words * W * nvox * outer / (ms_per_step * 1e-3) / 1e9
    print(json.dumps({"metric": "tvd_mvoxel_iter_per_s", "value": nvox * outer / (ms_per_step * 1e-3) / 1e6,
                      "unit": "Mvoxel*iter/s", "n_gpus": world, "ms_per_step": ms_per_step, "scaling": "strong",
                      "config": config_of(w, wname), "transport": "nccl (round-1 Python driver)",
                      "roofline": {"frac": gbs / world / peak, "peak": peak, "per_gpu": True}}))


def run_series(args, w, wname):
    """BASELINE configs[2]: a time series of frames, heterogeneous (per-pixel) weight / omega, frame-sharded over the
    ranks with no collective.  Strong scaling: the series has w["series"] frames in total, every rank takes its block.
    value: device-resident (the rank's frames solved in resident sub-batches of nb frames); e2e: host frames in, host
    frames out through the three-stage streaming pipeline (pageable frames, pinned rings, copies overlapped)."""
    import torch
    import torch.distributed as dist

    import darsia_b200 as da
    from darsia_b200 import _cabi
    from darsia_b200.engine import Engine, SolverSpec

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local])
        torch.cuda.synchronize(dev)

    def reduce_max(x):
        tx = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tx, op=dist.ReduceOp.MAX)
        return float(tx.item())

    T = int(args.frames or w["series"])
    s, e = da.frame_range(T, rank, world)
    mine = e - s
    B = w["nb"]
    shape = tuple(w["shape"])
    kind, m, tol = w["solver"]
    spec = SolverSpec(kind=_cabi.TVDK_SOLVER_JACOBI, maxiter=m, tol=tol)
    np_dtype = np.dtype(w["dtype"])
    outer = w["outer"]
    fvox = int(np.prod(shape))
    mu_h, om_h = hetero_maps(w, seed=3)
    pool = [phantom_np(shape, 3 + 100 * rank + k, w["dtype"]) for k in range(4)]  # distinct pageable host frames, cycled
    pool_pinned = [torch.from_numpy(a).pin_memory().numpy() for a in pool]  # the same frames page-locked (the e2e contract)
    eng = Engine(shape, nb=B, dtype=np_dtype, device=dev)
    mu_f, om_f = eng.coef(mu_h)[1], eng.coef(om_h)[1]
    f = eng.upload(np.stack([pool[k % 4] for k in range(B)], axis=-1))
    u, t = eng.empty_fields(1), eng.empty_fields(2)
    nsub = (mine + B - 1) // B

    def device_pass():
        for _ in range(nsub):
            u.copy_(f)
            t.zero_()
            eng.solve(f, u, t, mu_f, om_f, ELL, False, spec, outer, None, sync=False)

    for _ in range(max(1, args.warmup)):
        eng.solve(f, u.copy_(f), t.zero_(), mu_f, om_f, ELL, False, spec, outer, None, sync=False)
    lib = _cabi.get_library()
    lib.launch_count(1)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        device_pass()
    ev1.record()
    barrier()
    launches = int(lib.launch_count(1))
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = reduce_max(ev0.elapsed_time(ev1)) / args.steps
    # (every rank solved nsub full sub-batches; count the frames of the series, not the padding of the last one)
    value = T * fvox * outer / (ms_per_step * 1e-3) / 1e6
    W = np_dtype.itemsize
    words = algo_words_per_voxel_iter(2, w["solver"], 2)
    peak, peak_src = measured_peak()
    gbs = words * W * nsub * B * fvox * outer / (ms_per_step * 1e-3) / 1e9
    # ---- e2e: the streaming pipeline on pageable host frames ----
    kw = dict(dim=2, batch=B, mu=mu_h, omega=om_h, ell=ELL, isotropic=False, max_num_iter=outer, solver=da.Jacobi(maxiter=1),
              device=dev, rank=rank, world=world)
    got = {}

    def sink(k, fr):
        if k == s:
            got["first"] = fr.copy()

    da.tvd_series_stream(lambda k: pool[k % 4], n_frames=min(T, 2 * B * world), sink=lambda k, fr: None, **kw)  # allocator warm-up
    barrier()
    t0 = time.perf_counter()
    da.tvd_series_stream(lambda k: pool_pinned[k % 4], n_frames=T, sink=sink, **kw)  # frames in pinned host memory
    torch.cuda.synchronize(dev)
    te = reduce_max(time.perf_counter() - t0)
    barrier()
    t0 = time.perf_counter()
    da.tvd_series_stream(lambda k: pool[k % 4], n_frames=T, sink=lambda k, fr: None, **kw)  # pageable frames: + one staging memcpy each
    torch.cuda.synchronize(dev)
    te_pageable = reduce_max(time.perf_counter() - t0)
    # the streamed result of the first frame == the resident solve of the same frame
    chk = da.split_bregman_tvd(pool[s % 4], mu=mu_h, omega=om_h, ell=ELL, dim=2, max_num_iter=outer, solver=da.Jacobi(maxiter=1),
                               device=dev)
    same = bool(np.array_equal(chk, got["first"]))
    if world > 1:
        dist.destroy_process_group()
    if rank != 0:
        return
    line = {
        "metric": "tvd_mvoxel_iter_per_s", "value": value, "unit": "Mvoxel*iter/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64" if W == 8 else "f32", "data": "synthetic (4 distinct seeded frames per rank, cycled; per-pixel mu / omega maps)",
        "config": dict(config_of(w, wname), frames=T, frames_per_rank=mine, sub_batch=B,
                       sharding="contiguous blocks of frames per rank, no collective"),
        "roofline": {"bound": "hbm", "kernel": "whole step (fused Jacobi-1 iterations with per-pixel maps)", "achieved": gbs,
                     "peak": peak, "unit": "GB/s", "frac": gbs / peak, "traffic": None, "peak_source": peak_src,
                     "algo_words_per_voxel_iter": words, "per_gpu": True},
        "cpu_baseline": None,
        "e2e": {"value": T * fvox * outer / te / 1e6, "unit": "Mvoxel*iter/s", "s_per_step": te,
                "h2d_bytes_per_step": int(mine * fvox * W), "d2h_bytes_per_step": int(mine * fvox * W),
                "input": "frames in pinned host memory from a generator (tvd_series_stream: H2D / solve / D2H overlapped, results into pinned rings)",
                "matches_resident_result": same,
                "pageable_input": {"value": T * fvox * outer / te_pageable / 1e6, "s_per_step": te_pageable,
                                   "note": "pageable frames: one extra staging memcpy per frame on the host (a loader thread per rank)"}},
        "gpu_launches": launches, "clocks": clocks,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-slab", action="store_true", help="skip the slab-decomposed sub-record of the default workload")
    ap.add_argument("--transport", default="p2p", choices=["p2p", "nccl"], help="slab workloads: tvdk_solve_slab or the NCCL driver")
    ap.add_argument("--shape", type=int, nargs="+", default=None, help="override the workload's spatial shape")
    ap.add_argument("--frames", type=int, default=None, help="series workloads: total number of frames (default: the workload's)")
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    if args.shape:
        w = dict(w, shape=tuple(args.shape))
    if args.impl == "reference":
        run_reference_arm(args, w, args.workload)
    elif w.get("series"):
        run_series(args, w, args.workload)
    elif w.get("slab"):
        run_slab(args, w, args.workload)
    else:
        run_ours(args, w, args.workload)


if __name__ == "__main__":
    main()
