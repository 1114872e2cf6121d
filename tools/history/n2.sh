#!/bin/bash
mkdir -p gpurun_out
N=2
python bench.py --no-cpu --steps 3 > gpurun_out/bench_e2e_dbg.json 2> gpurun_out/bench_e2e_dbg.err
python -c "
import json; d=json.load(open('gpurun_out/bench_e2e_dbg.json')); print(d['value'], d['e2e'])"
python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/pytest_multi_n$N.log 2>&1; tail -3 gpurun_out/pytest_multi_n$N.log
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N "${@:2}" \
     > gpurun_out/$1_n$N.json 2> gpurun_out/$1_n$N.err; tail -c 300 gpurun_out/$1_n$N.json; }
run slab2_cfg4 --steps 2 --warmup 3 --workload cfg4_3d_slab_cg10 --shape 384 512 512
run slab2_cfg5_f32 --steps 2 --warmup 3 --workload cfg5_3d_slab_j1_f32 --shape 512 1024 512
