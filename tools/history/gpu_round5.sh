#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/pytest_parity_v5.log 2>&1; tail -3 gpurun_out/pytest_parity_v5.log
out=gpurun_out/sweepn.txt; : > $out
python tools/kbench.py --shape 4000 7000 --nb 3 --kernels sweep,sweep_x2,sweep_x4 --reps 30 >> $out 2>&1
python tools/kbench.py --shape 1788 3180 --nb 1 --kernels sweep,sweep_x2,sweep_x4 --reps 100 >> $out 2>&1
python tools/kbench.py --shape 512 512 512 --nb 1 --kernels sweep,sweep_x2 --reps 30 >> $out 2>&1
python tools/kbench.py --shape 4000 7000 --nb 3 --kernels sweep_x4 --tile-x 64,128,256 --zchunk 32,64,128 --reps 30 >> $out 2>&1
python tools/kbench.py --shape 512 512 512 --nb 1 --kernels sweep_x2 --tile-y 12,16 --zchunk 32,64 --reps 30 >> $out 2>&1
cat $out
python bench.py --no-cpu --workload cfg1_2d_aniso_j20 > gpurun_out/bench_cfg1_j20.json 2> gpurun_out/bench_cfg1_j20.err
python -c "
import json; d=json.load(open('gpurun_out/bench_cfg1_j20.json')); print(d['value'], d['ms_per_step'], d['roofline'], d['e2e']['value'])"
tail -3 gpurun_out/bench_cfg1_j20.err
