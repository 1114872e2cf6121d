#!/bin/bash
# L2-prefetch-distance sweep of the march kernels (TVDK_PF planes ahead; 0 = off)
mkdir -p gpurun_out
out=gpurun_out/pf_sweep.txt; : > $out
for pf in 0 2 4 8 16; do
  echo "== TVDK_PF=$pf" >> $out
  TVDK_PF=$pf python tools/kbench.py --shape 512 512 512 --nb 1 --kernels cg_update,cg_pq,fj1_aniso,fj1_iso,shrink,cg_init >> $out 2>&1
  TVDK_PF=$pf python tools/kbench.py --shape 512 512 512 --nb 1 --dtype float32 --kernels cg_update,fj1_aniso >> $out 2>&1
  TVDK_PF=$pf python tools/kbench.py --shape 4000 7000 --nb 3 --kernels cg_update,cg_pq,cg_init,shrink,fj1_aniso,fj1_iso,sweep >> $out 2>&1
done
for pf in 0 4; do
  echo "== TVDK_PF=$pf tile sweep 3d" >> $out
  TVDK_PF=$pf python tools/kbench.py --shape 512 512 512 --nb 1 --kernels cg_update,cg_pq,fj1_aniso --tile-y 8,16 --zchunk 16,32,64 >> $out 2>&1
done
python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/pytest_parity.log 2>&1; tail -3 gpurun_out/pytest_parity.log
cat $out
