#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_v4.log 2>&1; tail -3 gpurun_out/pytest_gpu_v4.log
python bench.py --no-cpu --workload cfg1_2d_aniso_j1 > gpurun_out/bench_cfg1_v2.json 2> gpurun_out/bench_cfg1_v2.err
python -c "
import json; d=json.load(open('gpurun_out/bench_cfg1_v2.json')); print(d['value'], d['e2e'])"
python bench.py --no-cpu > gpurun_out/bench_default_v4.json 2> gpurun_out/bench_default_v4.err
python -c "
import json; d=json.load(open('gpurun_out/bench_default_v4.json')); print(d['value'], d['ms_per_step'], d['e2e'])"
python tools/kbench.py --shape 4000 7000 --nb 3 --kernels cg_fused --tile-x 128,256 --zchunk 32,48,64 --reps 30 > gpurun_out/fused_tiles.txt 2>&1; cat gpurun_out/fused_tiles.txt
python __graft_entry__.py smoke
