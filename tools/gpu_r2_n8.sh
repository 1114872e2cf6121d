#!/bin/bash
# Round 2, 8-GPU session (run under gpurun --gpus 8): multi-GPU tests, default bench at N = 8 (incl. the slab sub-records),
# streamed series, one NCCL-transport run for comparison.
N=${1:-8}
set -x
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/r2n${N}_pytest_multi.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2n${N}_pytest_multi.log
tail -5 gpurun_out/r2n${N}_pytest_multi.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 1200 $TR bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2n${N}_bench_default.json 2> gpurun_out/r2n${N}_bench_default.err; echo "rc=$?"
tail -c 7000 gpurun_out/r2n${N}_bench_default.json; tail -5 gpurun_out/r2n${N}_bench_default.err
timeout 900 $TR bench.py --gpus $N --workload cfg3_series512 --steps 1 --warmup 1 > gpurun_out/r2n${N}_cfg3_series512.json 2> gpurun_out/r2n${N}_cfg3_series512.err; echo "rc=$?"
tail -c 2500 gpurun_out/r2n${N}_cfg3_series512.json; tail -3 gpurun_out/r2n${N}_cfg3_series512.err
timeout 600 $TR bench.py --gpus $N --workload cfg4_3d_slab_cg10 --steps 3 --warmup 2 --transport nccl > gpurun_out/r2n${N}_cfg4_nccl.json 2> gpurun_out/r2n${N}_cfg4_nccl.err; echo "rc=$?"
cat gpurun_out/r2n${N}_cfg4_nccl.json; tail -3 gpurun_out/r2n${N}_cfg4_nccl.err
timeout 600 $TR bench.py --gpus $N --workload cfg5_3d_slab_j1_f64 --steps 3 --warmup 2 > gpurun_out/r2n${N}_cfg5_f64.json 2> gpurun_out/r2n${N}_cfg5_f64.err; echo "rc=$?"
tail -c 1800 gpurun_out/r2n${N}_cfg5_f64.json; tail -3 gpurun_out/r2n${N}_cfg5_f64.err
