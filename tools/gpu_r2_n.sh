#!/bin/bash
# Round 2, multi-GPU session: bash tools/gpu_r2_n.sh N   (run under gpurun --gpus N)
N=${1:-2}
set -x
mkdir -p gpurun_out
nvidia-smi -L
nvidia-smi topo -m | head -12
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/r2n${N}_pytest_multi.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2n${N}_pytest_multi.log
tail -5 gpurun_out/r2n${N}_pytest_multi.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
for wl in cfg4_3d_slab_cg10 cfg5_3d_slab_j1_f32; do
  timeout 600 $TR bench.py --gpus $N --workload $wl --steps 3 --warmup 2 > gpurun_out/r2n${N}_${wl}_p2p.json 2> gpurun_out/r2n${N}_${wl}_p2p.err; echo "rc=$?"
  tail -c 2500 gpurun_out/r2n${N}_${wl}_p2p.json; tail -3 gpurun_out/r2n${N}_${wl}_p2p.err
  timeout 600 $TR bench.py --gpus $N --workload $wl --steps 3 --warmup 2 --transport nccl > gpurun_out/r2n${N}_${wl}_nccl.json 2> gpurun_out/r2n${N}_${wl}_nccl.err; echo "rc=$?"
  cat gpurun_out/r2n${N}_${wl}_nccl.json; tail -3 gpurun_out/r2n${N}_${wl}_nccl.err
done
timeout 900 $TR bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2n${N}_bench_default.json 2> gpurun_out/r2n${N}_bench_default.err; echo "rc=$?"
tail -c 6000 gpurun_out/r2n${N}_bench_default.json; tail -5 gpurun_out/r2n${N}_bench_default.err
