#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_final.log 2>&1; tail -3 gpurun_out/pytest_gpu_final.log
python __graft_entry__.py smoke 2>&1 | tail -4
python - <<'PY'
import time, numpy as np, torch, sys
sys.path.insert(0, '.')
import bench
from darsia_b200.restoration.chambolle import denoise_tv_chambolle
img = bench.phantom_np((4000, 7000), 3, "float64")
t = torch.from_numpy(img).cuda()
for _ in range(2):
    out, info = denoise_tv_chambolle(t, 0.1, 0.0, 50, return_info=True)
torch.cuda.synchronize(); t0 = time.perf_counter()
out, info = denoise_tv_chambolle(t, 0.1, 0.0, 50, return_info=True)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("chambolle 4000x7000 fp64: %d iterations in %.3f s -> %.0f Mvoxel*iter/s, %.0f GB/s algorithmic (5 words)" % (info["iterations"], dt, 28e6*info["iterations"]/dt/1e6, 5*8*28e6*info["iterations"]/dt/1e9))
PY
