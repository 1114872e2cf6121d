#!/bin/bash
mkdir -p gpurun_out
out=gpurun_out/tile3d.txt; : > $out
python tools/kbench.py --shape 512 512 512 --nb 1 --kernels cg_pq,cg_update,cg_init,shrink,sweep,fj1_aniso --tile-y 8,16 --zchunk 32 --reps 30 >> $out 2>&1
python tools/kbench.py --shape 512 512 512 --nb 1 --dtype float32 --kernels cg_pq,cg_update,cg_init,shrink,fj1_aniso --tile-y 8,16 --zchunk 32 --reps 30 >> $out 2>&1
echo "== sustained" >> $out
python tools/kbench.py --shape 4000 7000 --nb 3 --kernels cg_fused --reps 30 >> $out 2>&1
python tools/kbench.py --shape 4000 7000 --nb 3 --kernels cg_fused --reps 2000 >> $out 2>&1
python tools/kbench.py --shape 4000 7000 --nb 3 --kernels cg_update --reps 2000 >> $out 2>&1
python tools/kbench.py --shape 4000 7000 --nb 3 --kernels shrink --reps 2000 >> $out 2>&1
cat $out
python bench.py --no-cpu > gpurun_out/bench_default_v5.json 2> gpurun_out/bench_default_v5.err
python -c "
import json; d=json.load(open('gpurun_out/bench_default_v5.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'])"
