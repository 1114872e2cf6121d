#!/bin/bash
# One GPU session: parity tests, default bench, ncu launch list, ncu --set full of the top kernels.
set -x
mkdir -p gpurun_out
nvidia-smi -L
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"
python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/bench_short.json 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_launches.log 2>&1
python tools/kbench.py --shape 4000 7000 --nb 3 --kernels cg_update --reps 3 > gpurun_out/kb_a.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:tvdk_kernel --launch-skip 3 -c 1 -o gpurun_out/full_cg_update_2d -f \
    python tools/kbench.py --shape 4000 7000 --nb 3 --kernels cg_update --reps 3 > gpurun_out/ncu_full_a.log 2>&1
python tools/kbench.py --shape 512 512 512 --nb 1 --kernels fj1_aniso --reps 3 > gpurun_out/kb_b.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:tvdk_kernel --launch-skip 3 -c 1 -o gpurun_out/full_fj1_3d -f \
    python tools/kbench.py --shape 512 512 512 --nb 1 --kernels fj1_aniso --reps 3 > gpurun_out/ncu_full_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tvdk_kernel --launch-skip 3 -c 1 -o gpurun_out/full_cg_update_3d -f \
    python tools/kbench.py --shape 512 512 512 --nb 1 --kernels cg_update --reps 3 > gpurun_out/ncu_full_c.log 2>&1
ls -la gpurun_out
