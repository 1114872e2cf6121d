"""Where does a PAGEABLE input lose time against a pinned one?  (split_bregman_tvd's own timing split, 2 outer iterations)"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["DARSIA_B200_TIMING"] = "1"
import darsia_b200 as da  # noqa: E402
import importlib  # noqa: E402
from darsia_b200.engine import Engine  # noqa: E402

sbt = importlib.import_module("darsia_b200.restoration.split_bregman_tvd")
rng = np.random.default_rng(0)
img = rng.random((4000, 7000, 3))
pin = torch.from_numpy(img).pin_memory().numpy()
kw = dict(mu=0.1, omega=1.0, ell=1.0, dim=2, max_num_iter=2, isotropic=True, solver=da.CG(maxiter=10, tol=0.0))
for threads in (1, 4, 8):
    Engine.STAGE_THREADS = threads
    Engine._STAGE_POOL = None
    for name, arr in (("pinned", pin), ("pageable", img)):
        da.split_bregman_tvd(arr, **kw)
        torch.cuda.synchronize()
        ts = []
        for _ in range(3):
            t0 = time.perf_counter()
            out = da.split_bregman_tvd(arr, **kw)
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
        print(f"threads={threads} {name:8s} total {min(ts) * 1e3:7.1f} ms  " +
              "  ".join(f"{k}={v * 1e3:.1f}" for k, v in sbt.LAST_TIMINGS.items()), flush=True)
