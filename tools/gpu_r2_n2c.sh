#!/bin/bash
# Round 2, 2-GPU session: the default bench at N = 2 with full-size slab sub-records (what the driver's scaling run does),
# the two-devices-one-thread test, K1 vs K1x2 on the cfg1 image.
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
( time timeout 1200 $TR bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2n2c_bench_default.json 2> gpurun_out/r2n2c_bench_default.err ) 2> gpurun_out/r2n2c_time.txt; echo "rc=$?"
tail -c 2500 gpurun_out/r2n2c_bench_default.json; tail -5 gpurun_out/r2n2c_bench_default.err; cat gpurun_out/r2n2c_time.txt
timeout 300 python -m pytest tests/test_gpu_multi.py -x -q -k "two_devices" > gpurun_out/r2n2c_pytest_two_devices.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2n2c_pytest_two_devices.log
timeout 300 python tools/kbench.py --shape 1788 3180 --nb 1 --dtype float64 --kernels fj1_aniso,fj1x2_aniso,fj1_iso,sweep,sweep_x2,sweep_x4 --reps 50 2>&1 | grep -E "fj1|sweep" > gpurun_out/r2n2c_cfg1_kernels.txt; cat gpurun_out/r2n2c_cfg1_kernels.txt
