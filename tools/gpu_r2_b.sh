#!/bin/bash
# Round 2, session B (1 GPU): TMA-staged fused Jacobi-1 3-D kernel -- parity on the device, then A/B against the march kernel.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_slab.py tests/test_gpu_configs.py -x -q -k "tma or slabs or cfg or energy" > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
tail -5 gpurun_out/r2b_pytest.log
for dt in float64 float32; do
  for tma in 1 0; do
    TVDK_TMA=$tma timeout 300 python tools/kbench.py --shape 512 512 512 --nb 1 --dtype $dt --kernels fj1_aniso --reps 20 2>&1 | sed "s/^/TMA=$tma /"
  done
done > gpurun_out/r2b_tma_ab.txt 2>&1
# tile sweep of the TMA kernel
for dt in float64 float32; do
  TVDK_TMA=1 timeout 600 python tools/kbench.py --shape 512 512 512 --nb 1 --dtype $dt --kernels fj1_aniso --reps 10 --tile-x 16,32,64 --tile-y 8,12,16,32 --zchunk 32,64,128 2>&1
done > gpurun_out/r2b_tma_tiles.txt 2>&1
cat gpurun_out/r2b_tma_ab.txt
sort -k9 -n gpurun_out/r2b_tma_tiles.txt | tail -30
