#!/bin/bash
# Round 2, session H (1 GPU): GPU test-suite on the rebuilt library (MG per-voxel / warm start, explicit-state operators),
# host->device probe for pageable inputs, default bench with the in-step kernel timing.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2h_pytest.log
tail -4 gpurun_out/r2h_pytest.log
timeout 300 python tools/h2d_probe.py 672 > gpurun_out/r2h_h2d_probe.txt 2>&1; cat gpurun_out/r2h_h2d_probe.txt
timeout 900 python bench.py --gpus 1 --steps 5 --warmup 3 > gpurun_out/r2h_bench_default.json 2> gpurun_out/r2h_bench_default.err; echo "bench rc=$?"
tail -c 2500 gpurun_out/r2h_bench_default.json; tail -3 gpurun_out/r2h_bench_default.err
