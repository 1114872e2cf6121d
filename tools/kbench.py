#!/usr/bin/env python
"""Time individual kernels through the C ABI (CUDA events on the launching stream).

    python tools/kbench.py [--shape 4000 7000] [--nb 3] [--dtype float64] [--kernels cg_update,cg_pq,...]
                           [--tile-x 0,64,128,256] [--zchunk 0,32,64,128] [--reps 20]

Prints one line per (kernel, tile) with the average launch time, the algorithmic GB/s of the
implemented formulation and the fraction of the measured HBM peak.  Used for tuning and as the
short command that ncu profiles (tools/prof_driver.py wraps it).
"""
import argparse
import ctypes as C
import itertools
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", type=int, nargs="+", default=[4000, 7000])
    ap.add_argument("--nb", type=int, default=3)
    ap.add_argument("--dtype", default="float64")
    ap.add_argument("--kernels", default="cg_update,cg_pq,cg_init,shrink,fj1_iso,fj1_aniso")
    ap.add_argument("--tile-x", default="0")
    ap.add_argument("--tile-y", default="0")
    ap.add_argument("--zchunk", default="0")
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--hetero", action="store_true")
    args = ap.parse_args()

    import torch

    from darsia_b200 import _cabi
    from darsia_b200.engine import Engine

    peak = 6433.3
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = float(json.load(open(pk))["hbm_gbs"])
    dev = torch.device("cuda", 0)
    lib = _cabi.get_library()
    shape = tuple(args.shape)
    dim = len(shape)
    dt = np.dtype(args.dtype)
    eng = Engine(shape, nb=args.nb, dtype=dt, device=dev)
    g = eng.geom
    N = int(np.prod(shape)) * args.nb
    W = dt.itemsize
    gen = torch.Generator(device=dev).manual_seed(1)

    def rnd(n=1):
        return torch.rand(n * eng.batch_elems, dtype=eng.tdtype, device=dev, generator=gen)

    f, u, u2, x, r, p0, p1, r2 = rnd(), rnd(), rnd(), rnd(), rnd(), rnd(), rnd(), rnd()
    t, t2 = 0.1 * rnd(dim), rnd(dim)
    sums = torch.zeros(4 * args.nb, dtype=torch.float64, device=dev)
    coef = torch.full((args.nb,), 1e-3, dtype=torch.float64, device=dev)
    if args.hetero:
        mu_c, _k1 = eng.coef(0.1 * rnd()[: eng.field_elems].contiguous())
        om_c, _k2 = eng.coef(0.5 + rnd()[: eng.field_elems].contiguous())
        nh = 2
    else:
        mu_c, _ = eng.coef(0.1)
        om_c, _ = eng.coef(1.0)
        nh = 0
    el_c, _ = eng.coef(1.0)
    bg, bo, be, bm = C.byref(g), C.byref(om_c), C.byref(el_c), C.byref(mu_c)
    s = eng.stream()

    def ws():
        w_ = eng.op_workspace()
        return w_.data_ptr(), w_.numel()

    K = {
        "cg_update": (5, lambda: lib.cg_update(bg, bo, be, p0.data_ptr(), x.data_ptr(), r.data_ptr(), coef.data_ptr(),
                                               sums.data_ptr(), *ws(), s)),
        "cg_pq": (3, lambda: lib.cg_pq(bg, bo, be, p0.data_ptr(), r.data_ptr(), p1.data_ptr(), coef.data_ptr(), 0,
                                       sums.data_ptr(), *ws(), s)),
        "cg_fused": (6, lambda: lib.cg_fused(bg, bo, be, p0.data_ptr(), r.data_ptr(), p1.data_ptr(), r2.data_ptr(),
                                             x.data_ptr(), coef.data_ptr(), coef.data_ptr(), sums.data_ptr(), *ws(), s)),
        "cg_first": (1, lambda: lib.cg_first(bg, bo, be, r.data_ptr(), sums.data_ptr(), *ws(), s)),
        "cg_init": (4 + dim, lambda: lib.cg_init(bg, 1, bm, bo, be, f.data_ptr(), t.data_ptr(), eng.batch_elems,
                                                 u.data_ptr(), x.data_ptr(), r.data_ptr(), sums.data_ptr(), *ws(), s)),
        "shrink": (2 + 2 * dim, lambda: lib.shrink(bg, 1, bm, be, u.data_ptr(), u2.data_ptr(), None, t.data_ptr(),
                                                   t2.data_ptr(), eng.batch_elems, sums.data_ptr(), *ws(), s)),
        "sweep": (3, lambda: lib.jacobi_sweep(bg, bo, be, u.data_ptr(), f.data_ptr(), u2.data_ptr(), *ws(), s)),
        "sweep_x4": (3 / 4, lambda: lib.jacobi_sweeps(bg, bo, be, u.data_ptr(), f.data_ptr(), u2.data_ptr(), 4, None, *ws(), s)),
        "sweep_x2": (3 / 2, lambda: lib.jacobi_sweeps(bg, bo, be, u.data_ptr(), f.data_ptr(), u2.data_ptr(), 2, None, *ws(), s)),
        "matvec": (2, lambda: lib.cg_matvec(bg, bo, be, u.data_ptr(), u2.data_ptr(), *ws(), s)),
        "dot": (2, lambda: lib.dot(bg, u.data_ptr(), f.data_ptr(), sums.data_ptr(), *ws(), s)),
        "fj1x2_iso": ((3 + 2 * dim + nh) / 2, lambda: lib.fused_jacobi1_x2(bg, 1, bm, bo, be, f.data_ptr(), u.data_ptr(),
                                                                           u2.data_ptr(), t.data_ptr(), t2.data_ptr(),
                                                                           eng.batch_elems, sums.data_ptr(), *ws(), s)),
        "fj1x2_aniso": ((3 + 2 * dim + nh) / 2, lambda: lib.fused_jacobi1_x2(bg, 0, bm, bo, be, f.data_ptr(), u.data_ptr(),
                                                                             u2.data_ptr(), t.data_ptr(), t2.data_ptr(),
                                                                             eng.batch_elems, sums.data_ptr(), *ws(), s)),
        "fj1_iso": (3 + 2 * dim + nh, lambda: lib.fused_jacobi1(bg, 1, bm, bo, be, f.data_ptr(), u.data_ptr(),
                                                                u2.data_ptr(), t.data_ptr(), t2.data_ptr(),
                                                                eng.batch_elems, sums.data_ptr(), *ws(), s)),
        "fj1_aniso": (3 + 2 * dim + nh, lambda: lib.fused_jacobi1(bg, 0, bm, bo, be, f.data_ptr(), u.data_ptr(),
                                                                  u2.data_ptr(), t.data_ptr(), t2.data_ptr(),
                                                                  eng.batch_elems, sums.data_ptr(), *ws(), s)),
    }
    tiles = itertools.product([int(v) for v in args.tile_x.split(",")], [int(v) for v in args.tile_y.split(",")],
                              [int(v) for v in args.zchunk.split(",")])
    for tx, ty, zc in tiles:
        lib.check(lib.set_tune(tx, ty, zc), "set_tune")
        eng._op_ws = None
        for name in args.kernels.split(","):
            words, fn = K[name]
            for _ in range(2):
                lib.check(fn(), name)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.reps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.reps
            nsw = {"sweep_x4": 4, "sweep_x2": 2, "fj1x2_iso": 2, "fj1x2_aniso": 2}.get(name, 1)  # per sweep / iteration
            gbs = words * nsw * W * N / (ms * 1e-3) / 1e9
            print(f"{name:10s} shape={shape} nb={args.nb} {dt.name} tile=({tx},{ty},{zc}) {ms / nsw:8.4f} ms  "
                  f"{gbs:7.1f} GB/s algo ({words} words)  {gbs / peak:5.3f} of measured peak", flush=True)
    lib.set_tune(0, 0, 0)


if __name__ == "__main__":
    main()
