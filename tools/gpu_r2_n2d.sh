#!/bin/bash
# Round 2, 2-GPU session D: the default bench at N = 2 after the IPC-mapping leak fix (what the driver's scaling run does).
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
( time timeout 1200 $TR bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2n2d_bench_default.json 2> gpurun_out/r2n2d_bench_default.err ) 2> gpurun_out/r2n2d_time.txt; echo "rc=$?"
tail -c 3000 gpurun_out/r2n2d_bench_default.json; tail -5 gpurun_out/r2n2d_bench_default.err; cat gpurun_out/r2n2d_time.txt
nvidia-smi --query-gpu=memory.used --format=csv
