#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_v6.log 2>&1; tail -3 gpurun_out/pytest_gpu_v6.log
out=gpurun_out/kernels_v6.txt; : > $out
K2="cg_fused,cg_first,cg_update,cg_pq,cg_init,shrink,fj1_aniso,fj1_iso,sweep,sweep_x2,sweep_x4"
python tools/kbench.py --shape 4000 7000 --nb 3 --kernels $K2 --reps 30 >> $out 2>&1
python tools/kbench.py --shape 1788 3180 --nb 1 --kernels fj1_aniso,sweep,sweep_x2,sweep_x4,cg_fused --reps 100 >> $out 2>&1
python tools/kbench.py --shape 512 512 512 --nb 1 --kernels cg_update,cg_pq,cg_init,shrink,fj1_aniso,fj1_iso,sweep --reps 30 >> $out 2>&1
python tools/kbench.py --shape 4000 7000 --nb 8 --hetero --kernels fj1_aniso,fj1_iso --reps 20 >> $out 2>&1
cat $out
for w in cfg1_2d_aniso_j20 cfg1_2d_aniso_j1 cfg3_2d_hetero_j1; do
python bench.py --no-cpu --workload $w > gpurun_out/bench_${w}_v6.json 2> gpurun_out/bench_${w}_v6.err
python -c "
import json; d=json.load(open('gpurun_out/bench_${w}_v6.json')); print('$w', d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['whole_step']['frac'], d['e2e']['value'])"
done
python bench.py > gpurun_out/bench_default_v6.json 2> gpurun_out/bench_default_v6.err
python -c "
import json; d=json.load(open('gpurun_out/bench_default_v6.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['cpu_baseline']['value'])"
