#!/bin/bash
# L2-prefetch sweep of the march kernels: TVDK_PF planes ahead (0 = off), TVDK_PFSUB = one thread per line
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/pytest_parity.log 2>&1; tail -3 gpurun_out/pytest_parity.log
out=gpurun_out/pf_sweep2.txt; : > $out
K2="cg_fused,cg_first,cg_update,cg_pq,cg_init,shrink,fj1_aniso,fj1_iso,sweep"
for sub in 0 1; do
for pf in 0 1 2 3 4 6; do
  if [ $pf = 0 ] && [ $sub = 1 ]; then continue; fi
  echo "== PF=$pf SUB=$sub" >> $out
  TVDK_PF=$pf TVDK_PFSUB=$sub python tools/kbench.py --shape 512 512 512 --nb 1 --kernels $K2 >> $out 2>&1
  TVDK_PF=$pf TVDK_PFSUB=$sub python tools/kbench.py --shape 512 512 512 --nb 1 --dtype float32 --kernels cg_fused,cg_update,fj1_aniso >> $out 2>&1
  TVDK_PF=$pf TVDK_PFSUB=$sub python tools/kbench.py --shape 4000 7000 --nb 3 --kernels $K2 >> $out 2>&1
done
done
echo "== defaults" >> $out
python tools/kbench.py --shape 512 512 512 --nb 1 --kernels $K2 >> $out 2>&1
python tools/kbench.py --shape 4000 7000 --nb 3 --kernels $K2 >> $out 2>&1
python tools/kbench.py --shape 4000 7000 --nb 3 --kernels cg_fused --tile-x 64,128,256 --zchunk 16,32,64,128 >> $out 2>&1
python tools/kbench.py --shape 512 512 512 --nb 1 --kernels cg_fused --tile-y 8,12,16 --zchunk 16,32,64 >> $out 2>&1
python bench.py --no-cpu > gpurun_out/bench_fusedcg.json 2> gpurun_out/bench_fusedcg.err
cat gpurun_out/bench_fusedcg.json | cut -c1-1500
