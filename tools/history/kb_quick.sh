#!/bin/bash
mkdir -p gpurun_out
out=gpurun_out/kb_quick.txt; : > $out
K2="cg_fused,cg_first,cg_update,cg_pq,cg_init,shrink,fj1_aniso,fj1_iso,sweep"
python tools/kbench.py --shape 512 512 512 --nb 1 --kernels $K2 >> $out 2>&1
python tools/kbench.py --shape 512 512 512 --nb 1 --dtype float32 --kernels cg_update,fj1_aniso,fj1_iso >> $out 2>&1
python tools/kbench.py --shape 4000 7000 --nb 3 --kernels $K2 >> $out 2>&1
python tools/kbench.py --shape 4000 7000 --nb 8 --hetero --kernels fj1_aniso,fj1_iso >> $out 2>&1
python tools/kbench.py --shape 1788 3180 --nb 1 --kernels fj1_aniso,fj1_iso,cg_fused >> $out 2>&1
python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/pytest_parity.log 2>&1; tail -3 gpurun_out/pytest_parity.log
python bench.py --no-cpu > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
cat $out
