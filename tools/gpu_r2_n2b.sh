#!/bin/bash
# quick 2-GPU check of the bench code paths added late (full-size parity gather, pinned e2e inputs, streamed series)
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus 2 --workload cfg5_3d_slab_j1_f32 --shape 256 512 256 --steps 2 --warmup 1 > gpurun_out/r2n2b_cfg5.json 2> gpurun_out/r2n2b_cfg5.err; echo "rc=$?"
tail -c 1800 gpurun_out/r2n2b_cfg5.json; tail -5 gpurun_out/r2n2b_cfg5.err
timeout 600 $TR bench.py --gpus 2 --workload cfg4_3d_slab_cg10 --shape 256 256 256 --steps 2 --warmup 1 > gpurun_out/r2n2b_cfg4.json 2> gpurun_out/r2n2b_cfg4.err; echo "rc=$?"
tail -c 1800 gpurun_out/r2n2b_cfg4.json; tail -5 gpurun_out/r2n2b_cfg4.err
timeout 600 $TR bench.py --gpus 2 --workload cfg3_series512 --frames 48 --steps 1 --warmup 1 > gpurun_out/r2n2b_cfg3.json 2> gpurun_out/r2n2b_cfg3.err; echo "rc=$?"
tail -c 1500 gpurun_out/r2n2b_cfg3.json; tail -5 gpurun_out/r2n2b_cfg3.err
