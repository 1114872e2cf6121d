#!/bin/bash
# Round 2, second 8-GPU session: default bench at N = 8 with the full-size parity check and pinned e2e inputs, streamed series.
set -x
mkdir -p gpurun_out
N=8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 1200 $TR bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2n8b_bench_default.json 2> gpurun_out/r2n8b_bench_default.err; echo "rc=$?"
tail -c 3000 gpurun_out/r2n8b_bench_default.json; tail -5 gpurun_out/r2n8b_bench_default.err
timeout 900 $TR bench.py --gpus $N --workload cfg3_series512 --steps 1 --warmup 1 > gpurun_out/r2n8b_cfg3_series512.json 2> gpurun_out/r2n8b_cfg3_series512.err; echo "rc=$?"
tail -c 1500 gpurun_out/r2n8b_cfg3_series512.json; tail -3 gpurun_out/r2n8b_cfg3_series512.err
