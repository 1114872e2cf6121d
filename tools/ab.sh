#!/bin/bash
# A/B of two builds of libtvdk (default vs TVDK_LIB=alt), interleaved
mkdir -p gpurun_out
out=gpurun_out/ab4.txt; : > $out
ALT=$PWD/darsia_b200/csrc/libtvdk_alt.so
for r in 1 2; do
  for lib in "" "$ALT"; do
    echo "== round $r lib=${lib:-default}" >> $out
    TVDK_LIB=$lib python tools/kbench.py --shape 4000 7000 --nb 3 --kernels cg_fused,cg_init,shrink,cg_first,fj1_aniso,sweep_x4 --reps 30 >> $out 2>&1
    TVDK_LIB=$lib python tools/kbench.py --shape 512 512 512 --nb 1 --kernels fj1_aniso,cg_update,cg_pq,cg_init,shrink --reps 30 >> $out 2>&1
    TVDK_LIB=$lib python tools/kbench.py --shape 4000 7000 --nb 8 --hetero --kernels fj1_aniso --reps 20 >> $out 2>&1
  done
done
cat $out
