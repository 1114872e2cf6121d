#!/bin/bash
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_default_v3.json 2> gpurun_out/bench_default_v3.err; echo "bench rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_bench_v3.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_launches_v3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tvdk_kernel --launch-skip 3 -c 1 -o gpurun_out/full_cg_fused_2d -f \
    python tools/kbench.py --shape 4000 7000 --nb 3 --kernels cg_fused --reps 3 > gpurun_out/ncu_full_d.log 2>&1
python bench.py --no-cpu --workload cfg3_2d_hetero_j1 > gpurun_out/bench_cfg3_n1.json 2> gpurun_out/bench_cfg3_n1.err
python bench.py --no-cpu --workload cfg1_2d_aniso_j1 > gpurun_out/bench_cfg1_n1.json 2> gpurun_out/bench_cfg1_n1.err
python bench.py --steps 1 --warmup 3 --workload cfg4_3d_slab_cg10 > gpurun_out/slab_cfg4_n1.json 2> gpurun_out/slab_cfg4_n1.err
python bench.py --steps 1 --warmup 3 --workload cfg5_3d_slab_j1_f32 > gpurun_out/slab_cfg5_f32_n1.json 2> gpurun_out/slab_cfg5_f32_n1.err
out=gpurun_out/small_sweep.txt; : > $out
python tools/kbench.py --shape 1788 3180 --nb 1 --kernels fj1_aniso,cg_fused,cg_update,shrink --tile-x 64,128 --zchunk 8,12,16,24,32,48 --reps 50 >> $out 2>&1
tail -c 400 gpurun_out/*_n1.json
