#!/bin/bash
# Round 2, session C (1 GPU): TMA kernel after the edge fix + new default tiles: tests, timings, ncu capture.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
tail -4 gpurun_out/r2c_pytest.log
for dt in float64 float32; do
  for tma in 1 0; do
    TVDK_TMA=$tma timeout 300 python tools/kbench.py --shape 512 512 512 --nb 1 --dtype $dt --kernels fj1_aniso --reps 20 2>&1 | sed "s/^/TMA=$tma /"
  done
  TVDK_TMA=1 timeout 300 python tools/kbench.py --shape 256 2048 512 --nb 1 --dtype $dt --kernels fj1_aniso --reps 20 2>&1 | sed "s/^/TMA=1 /"
  TVDK_TMA=1 timeout 300 python tools/kbench.py --shape 512 512 512 --nb 1 --dtype $dt --kernels fj1_aniso --reps 10 --tile-x 8,16 --tile-y 4,6,8,10,12 --zchunk 64 2>&1
done > gpurun_out/r2c_tma.txt 2>&1
cat gpurun_out/r2c_tma.txt | grep fj1
timeout 600 python bench.py --workload cfg5_3d_slab_j1_f32 --steps 3 --warmup 2 > gpurun_out/r2c_bench_cfg5_f32_n1.json 2> gpurun_out/r2c_bench_cfg5.err; tail -c 1500 gpurun_out/r2c_bench_cfg5_f32_n1.json
# ncu: the TMA kernel, one launch, full set (after the same command exited 0 without ncu)
python tools/kbench.py --shape 512 512 512 --nb 1 --dtype float64 --kernels fj1_aniso --reps 3 > gpurun_out/r2c_kb_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:tvdk_kernel_tma --launch-skip 3 -c 1 -o gpurun_out/r2_full_fj1_tma_3d -f \
    python tools/kbench.py --shape 512 512 512 --nb 1 --dtype float64 --kernels fj1_aniso --reps 3 > gpurun_out/r2c_ncu.log 2>&1
tail -3 gpurun_out/r2c_ncu.log
ls -la gpurun_out | tail -5
