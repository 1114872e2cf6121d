#!/bin/bash
# 8-GPU validation: multi-GPU tests, frame-sharded bench (weak scaling), slab-decomposed volumes (strong scaling)
mkdir -p gpurun_out
N=${1:-8}
nvidia-smi -L | wc -l
python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/pytest_multi_n$N.log 2>&1; tail -3 gpurun_out/pytest_multi_n$N.log
run() { # name, extra args
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N "${@:2}" \
     > gpurun_out/$1_n$N.json 2> gpurun_out/$1_n$N.err
  tail -c 600 gpurun_out/$1_n$N.json
}
run bench_cfg2 --steps 2 --warmup 3 --no-cpu
run bench_cfg3 --steps 2 --warmup 3 --no-cpu --workload cfg3_2d_hetero_j1
run slab_cfg4 --steps 1 --warmup 3 --workload cfg4_3d_slab_cg10
run slab_cfg5_f32 --steps 1 --warmup 3 --workload cfg5_3d_slab_j1_f32
run slab_cfg5_f64 --steps 1 --warmup 3 --workload cfg5_3d_slab_j1_f64
