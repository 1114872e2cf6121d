#!/bin/bash
# Round 2, session A (1 GPU): full GPU test-suite, default bench (incl. the N=1 slab sub-record), short reference arm.
set -x
mkdir -p gpurun_out
nvidia-smi -L
nproc; free -g | head -2
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest_gpu.log
tail -5 gpurun_out/r2a_pytest_gpu.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2a_bench_default.json 2> gpurun_out/r2a_bench_default.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/r2a_bench_default.json; tail -5 gpurun_out/r2a_bench_default.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2a_bench_reference.json 2> gpurun_out/r2a_bench_reference.err; echo "ref rc=$?"
cat gpurun_out/r2a_bench_reference.json; tail -3 gpurun_out/r2a_bench_reference.err
