#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_v9.log 2>&1; tail -3 gpurun_out/pytest_gpu_v9.log
out=gpurun_out/tile3d_v2.txt; : > $out
python tools/kbench.py --shape 512 512 512 --nb 1 --kernels fj1_aniso,cg_pq,cg_init,shrink --tile-x 16,32,64 --tile-y 8,16,32 --zchunk 32 --reps 20 >> $out 2>&1
python tools/kbench.py --shape 512 512 512 --nb 1 --dtype float32 --kernels fj1_aniso --tile-x 16,32,64 --tile-y 8,16,32 --zchunk 32 --reps 20 >> $out 2>&1
python tools/kbench.py --shape 512 512 512 --nb 1 --kernels fj1_aniso,cg_pq,cg_init --tile-x 32 --tile-y 16 --zchunk 16,32,64,128 --reps 20 >> $out 2>&1
cat $out
