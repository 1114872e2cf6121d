#!/bin/bash
# Round 2, session K (1 GPU, the last GPU minutes): the 2-D TMA-staged fused Jacobi-1 kernel (KFusedJ1T<T, false, 2>):
# parity on the GPU, A/B against the march kernel on the cfg1 image (kernel alone and the whole cfg1 solve), then the whole
# GPU test-suite with the route switched on.
set -x
mkdir -p gpurun_out
timeout 90 python -m pytest tests/test_gpu_parity.py -x -q -k "tma_staged_fused_jacobi1_2d" > gpurun_out/r2k_pytest_tma2d.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2k_pytest_tma2d.log
tail -3 gpurun_out/r2k_pytest_tma2d.log
{
TVDK_TMA2D=0 timeout 60 python tools/kbench.py --shape 1788 3180 --nb 1 --dtype float64 --kernels fj1_aniso --reps 50 2>&1 | grep fj1 | sed 's/^/march  /'
TVDK_TMA2D=1 timeout 60 python tools/kbench.py --shape 1788 3180 --nb 1 --dtype float64 --kernels fj1_aniso --tile-x 0,32,96 --zchunk 0,16,32 --reps 50 2>&1 | grep fj1 | sed 's/^/tma2d  /'
TVDK_TMA2D=0 timeout 60 python tools/kbench.py --shape 4000 7000 --nb 3 --dtype float64 --kernels fj1_aniso --reps 20 2>&1 | grep fj1 | sed 's/^/march  /'
TVDK_TMA2D=1 timeout 60 python tools/kbench.py --shape 4000 7000 --nb 3 --dtype float64 --kernels fj1_aniso --tile-x 0,96 --reps 20 2>&1 | grep fj1 | sed 's/^/tma2d  /'
TVDK_TMA2D=0 timeout 60 python tools/kbench.py --shape 1788 3180 --nb 1 --dtype float32 --kernels fj1_aniso --reps 50 2>&1 | grep fj1 | sed 's/^/march  /'
TVDK_TMA2D=1 timeout 60 python tools/kbench.py --shape 1788 3180 --nb 1 --dtype float32 --kernels fj1_aniso --reps 50 2>&1 | grep fj1 | sed 's/^/tma2d  /'
} > gpurun_out/r2k_kbench_tma2d.txt; cat gpurun_out/r2k_kbench_tma2d.txt
for t in 0 1; do
  TVDK_TMA2D=$t timeout 90 python bench.py --workload cfg1_2d_aniso_j1 --steps 10 --warmup 3 --no-cpu --no-slab > gpurun_out/r2k_cfg1_j1_tma2d$t.json 2> gpurun_out/r2k_cfg1_j1_tma2d$t.err
  python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r2k_cfg1_j1_tma2d$t.json") if l.startswith("{")][-1])
print("cfg1_j1 TVDK_TMA2D=$t value", round(d["value"],1), "ms_per_step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), "whole_step_frac", round(d["roofline"]["whole_step"]["frac"],4), "parity", d["parity"])
PY
done
TVDK_TMA2D=1 timeout 200 python -m pytest tests -m gpu -x -q > gpurun_out/r2k_pytest_all_tma2d1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2k_pytest_all_tma2d1.log
tail -3 gpurun_out/r2k_pytest_all_tma2d1.log
