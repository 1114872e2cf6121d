#!/bin/bash
# Round 2, session G (1 GPU): the end-of-round measurement set -- full GPU test-suite, smoke, default bench and reference arm
# with the driver's flags, ncu launch list of the bench command, ncu --set full of the dominant kernel (K7).
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2g_pytest.log
tail -4 gpurun_out/r2g_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2g_smoke.log 2>&1; echo "smoke rc=$?"; tail -8 gpurun_out/r2g_smoke.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2g_bench_default.json 2> gpurun_out/r2g_bench_default.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/r2g_bench_default.json; tail -3 gpurun_out/r2g_bench_default.err
( time timeout 1200 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2g_bench_reference.json 2> gpurun_out/r2g_bench_reference.err ) 2> gpurun_out/r2g_reference_time.txt; echo "ref rc=$?"
cat gpurun_out/r2g_bench_reference.json; cat gpurun_out/r2g_reference_time.txt
python bench.py --steps 1 --warmup 3 --no-cpu --no-slab > gpurun_out/r2g_bench_short.json 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/r2g_launches_bench.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu --no-slab > gpurun_out/r2g_ncu_launches.log 2>&1
python tools/kbench.py --shape 4000 7000 --nb 3 --kernels cg_fused --reps 3 > gpurun_out/r2g_kb_k7.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:tvdk_kernel --launch-skip 3 -c 1 -o gpurun_out/r2_full_cg_fused_2d -f \
    python tools/kbench.py --shape 4000 7000 --nb 3 --kernels cg_fused --reps 3 > gpurun_out/r2g_ncu_k7.log 2>&1
tail -2 gpurun_out/r2g_ncu_k7.log; ls -la gpurun_out | grep r2g | head -20
