#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_v10.log 2>&1; tail -3 gpurun_out/pytest_gpu_v10.log
out=gpurun_out/k1_3d_v10.txt; : > $out
python tools/kbench.py --shape 512 512 512 --nb 1 --kernels fj1_aniso,fj1_iso --reps 30 >> $out 2>&1
python tools/kbench.py --shape 512 512 512 --nb 1 --dtype float32 --kernels fj1_aniso,fj1_iso --reps 30 >> $out 2>&1
python tools/kbench.py --shape 512 512 512 --nb 1 --kernels fj1_aniso --tile-x 32 --tile-y 16 --zchunk 32 --reps 30 >> $out 2>&1
python tools/kbench.py --shape 512 512 512 --nb 1 --dtype float32 --kernels fj1_aniso --tile-x 32 --tile-y 16 --zchunk 32 --reps 30 >> $out 2>&1
cat $out
python bench.py --no-cpu --workload cfg5_3d_aniso_j1_f32 > gpurun_out/bench_cfg5_f32_512.json 2> gpurun_out/bench_cfg5_f32_512.err
python -c "
import json; d=json.load(open('gpurun_out/bench_cfg5_f32_512.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'])"
