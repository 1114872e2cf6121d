#!/bin/bash
# Round 2, session E (1 GPU): isotropic TMA variant (parity + A/B), small-image chunk sweep, compute-sanitizer racecheck.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py tests/test_gpu_slab.py -x -q > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2e_pytest.log
tail -4 gpurun_out/r2e_pytest.log
for dt in float64 float32; do
  for tma in 1 0; do
    TVDK_TMA=$tma timeout 300 python tools/kbench.py --shape 512 512 512 --nb 1 --dtype $dt --kernels fj1_iso,fj1_aniso --reps 20 2>&1 | grep fj1 | sed "s/^/TMA=$tma /"
  done
  TVDK_TMA=1 timeout 300 python tools/kbench.py --shape 512 512 512 --nb 1 --dtype $dt --kernels fj1_iso --reps 10 --tile-x 16 --tile-y 6,8,10,12,16 --zchunk 64 2>&1 | grep fj1
done > gpurun_out/r2e_tma_iso.txt 2>&1
cat gpurun_out/r2e_tma_iso.txt
timeout 600 python tools/kbench.py --shape 1788 3180 --nb 1 --dtype float64 --kernels fj1_aniso,sweep --reps 50 --tile-x 64,128,256 --zchunk 6,8,10,12,14,16,20,24,32 2>&1 | grep -E "fj1|sweep" > gpurun_out/r2e_cfg1_chunks.txt
sort -k8 -n gpurun_out/r2e_cfg1_chunks.txt | head -12
timeout 600 python tools/sanitize_small.py > gpurun_out/r2e_sanitize_plain.log 2>&1 && \
timeout 1200 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitize_small.py > gpurun_out/r2e_racecheck.log 2>&1; echo "racecheck rc=$?"
tail -15 gpurun_out/r2e_racecheck.log
