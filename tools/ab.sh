#!/bin/bash
# A/B of two builds of libtvdk (default vs TVDK_LIB=alt), interleaved, 3 rounds
mkdir -p gpurun_out
out=gpurun_out/ab2.txt; : > $out
ALT=$PWD/darsia_b200/csrc/libtvdk_alt.so
for r in 1 2 3; do
  for lib in "" "$ALT"; do
    echo "== round $r lib=${lib:-default}" >> $out
    TVDK_LIB=$lib python tools/kbench.py --shape 4000 7000 --nb 3 --kernels cg_fused,cg_first,shrink,cg_init,fj1_aniso --reps 40 >> $out 2>&1
    TVDK_LIB=$lib python tools/kbench.py --shape 1788 3180 --nb 1 --kernels cg_fused,fj1_aniso,cg_update --reps 100 >> $out 2>&1
    TVDK_LIB=$lib python tools/kbench.py --shape 512 512 512 --nb 1 --kernels fj1_aniso,cg_update,cg_pq --reps 40 >> $out 2>&1
  done
done
cat $out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_v3.log 2>&1; tail -3 gpurun_out/pytest_gpu_v3.log
