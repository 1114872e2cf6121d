#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -k "jacobi or golden" > gpurun_out/pytest_parity_v7.log 2>&1; tail -3 gpurun_out/pytest_parity_v7.log
out=gpurun_out/sweepn_v7.txt; : > $out
for r in 1 2; do
python tools/kbench.py --shape 4000 7000 --nb 3 --kernels sweep,sweep_x2,sweep_x4 --reps 30 >> $out 2>&1
python tools/kbench.py --shape 1788 3180 --nb 1 --kernels sweep,sweep_x2,sweep_x4 --reps 100 >> $out 2>&1
done
cat $out
python bench.py --no-cpu --workload cfg1_2d_aniso_j20 > gpurun_out/bench_cfg1_j20_v7.json 2> gpurun_out/bench_cfg1_j20_v7.err
python -c "
import json; d=json.load(open('gpurun_out/bench_cfg1_j20_v7.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'])"
