"""How fast can a PAGEABLE host array reach the device on this box?  (decides Engine._stage_pageable's defaults)
python tools/h2d_probe.py [MB]"""
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

mb = int(sys.argv[1]) if len(sys.argv) > 1 else 672
n = mb * (1 << 20) // 8
src = np.random.default_rng(0).random(n)
dev = torch.device("cuda", 0)
out = torch.empty(n, dtype=torch.float64, device=dev)
torch.cuda.synchronize()


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return src.nbytes / ((time.perf_counter() - t) / reps) / 1e9


print(f"{mb} MB fp64, host cpus {torch.get_num_threads()} torch threads")
print(f"direct copy_ from pageable       : {timed(lambda: out.copy_(torch.from_numpy(src))):6.2f} GB/s")
pin = torch.empty(n, dtype=torch.float64, pin_memory=True)
pin.numpy()[:] = src
print(f"copy_ from pinned                : {timed(lambda: out.copy_(pin, non_blocking=True)):6.2f} GB/s")
for chunk_mb in (4, 16, 64):
    for nt in (1, 2, 4, 8):
        chunk = chunk_mb * (1 << 20) // 8
        bufs = [torch.empty(chunk, dtype=torch.float64, pin_memory=True) for _ in range(2)]
        pool = ThreadPoolExecutor(nt) if nt > 1 else None

        def staged():
            evs = [None, None]
            for i, off in enumerate(range(0, n, chunk)):
                m = min(chunk, n - off)
                b = i & 1
                if evs[b] is not None:
                    evs[b].synchronize()
                dst = bufs[b].numpy()
                if pool is None:
                    dst[:m] = src[off:off + m]
                else:
                    step = (m + nt - 1) // nt
                    list(pool.map(lambda k: np.copyto(dst[k * step:min(m, (k + 1) * step)], src[off + k * step:off + min(m, (k + 1) * step)]),
                                  range(nt)))
                out[off:off + m].copy_(bufs[b][:m], non_blocking=True)
                evs[b] = torch.cuda.Event()
                evs[b].record()

        print(f"staged, {chunk_mb:3d} MB chunks, {nt} thread(s): {timed(staged):6.2f} GB/s")
t = time.perf_counter()
rc = torch.cuda.cudart().cudaHostRegister(src.ctypes.data, src.nbytes, 0)
t_reg = time.perf_counter() - t
reg = torch.from_numpy(src)
print(f"cudaHostRegister rc={rc}: {t_reg * 1e3:.1f} ms; copy_ from the registered array: {timed(lambda: out.copy_(reg, non_blocking=True)):6.2f} GB/s")
t = time.perf_counter()
torch.cuda.cudart().cudaHostUnregister(src.ctypes.data)
print(f"cudaHostUnregister: {(time.perf_counter() - t) * 1e3:.1f} ms")
