#!/bin/bash
# Round 2, 2-GPU session E: the multi-GPU tests and the default bench at N = 2 with programmatic dependent launch on
# (the build's default now) -- slab kernels behind IPC flags, full-size parity sub-records.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/r2n2e_pytest_multi.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2n2e_pytest_multi.log
tail -4 gpurun_out/r2n2e_pytest_multi.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
( time timeout 900 $TR bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2n2e_bench_default.json 2> gpurun_out/r2n2e_bench_default.err ) 2> gpurun_out/r2n2e_time.txt; echo "rc=$?"
tail -c 3000 gpurun_out/r2n2e_bench_default.json; tail -5 gpurun_out/r2n2e_bench_default.err; cat gpurun_out/r2n2e_time.txt
