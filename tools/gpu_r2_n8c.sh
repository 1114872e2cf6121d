#!/bin/bash
# Round 2, third (short) 8-GPU session: streamed series after the D2H-overlap fix; two-devices-one-thread test.
set -x
mkdir -p gpurun_out
N=8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 900 $TR bench.py --gpus $N --workload cfg3_series512 --steps 1 --warmup 1 > gpurun_out/r2n8c_cfg3_series512.json 2> gpurun_out/r2n8c_cfg3_series512.err; echo "rc=$?"
tail -c 1500 gpurun_out/r2n8c_cfg3_series512.json; tail -3 gpurun_out/r2n8c_cfg3_series512.err
timeout 300 python -m pytest tests/test_gpu_multi.py -x -q -k "two_devices" > gpurun_out/r2n8c_pytest_two_devices.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2n8c_pytest_two_devices.log
