#!/bin/bash
# Round 2, session J (1 GPU): the end-of-round measurement set on the final library (programmatic dependent launch on,
# MG per-voxel coefficients, in-step timing): GPU test-suite, smoke, default bench with the driver's flags, ncu launch list
# of the bench command, ncu --set full of K7 (cfg2) and of the 2-D fused Jacobi-1 kernel on the cfg1 image.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2j_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2j_pytest.log
tail -4 gpurun_out/r2j_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2j_smoke.log 2>&1; echo "smoke rc=$?"; tail -8 gpurun_out/r2j_smoke.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2j_bench_default.json 2> gpurun_out/r2j_bench_default.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/r2j_bench_default.json; tail -3 gpurun_out/r2j_bench_default.err
python bench.py --steps 1 --warmup 3 --no-cpu --no-slab > gpurun_out/r2j_bench_short.json 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/r2j_launches_bench.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu --no-slab > gpurun_out/r2j_ncu_launches.log 2>&1
python tools/kbench.py --shape 4000 7000 --nb 3 --kernels cg_fused --reps 3 > gpurun_out/r2j_kb_k7.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:tvdk_kernel --launch-skip 3 -c 1 -o gpurun_out/r2j_full_cg_fused_2d -f \
    python tools/kbench.py --shape 4000 7000 --nb 3 --kernels cg_fused --reps 3 > gpurun_out/r2j_ncu_k7.log 2>&1
python tools/kbench.py --shape 1788 3180 --nb 1 --kernels fj1_aniso --reps 3 > gpurun_out/r2j_kb_fj1_cfg1.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:tvdk_kernel --launch-skip 3 -c 1 -o gpurun_out/r2j_full_fj1_cfg1 -f \
    python tools/kbench.py --shape 1788 3180 --nb 1 --kernels fj1_aniso --reps 3 > gpurun_out/r2j_ncu_fj1_cfg1.log 2>&1
tail -2 gpurun_out/r2j_ncu_k7.log gpurun_out/r2j_ncu_fj1_cfg1.log; ls -la gpurun_out | grep r2j | head -20
