#!/bin/bash
# Final single-GPU measurements: default bench, reference arm, ncu launch list, ncu --set full of the top kernels
# (summarised on the box: the .ncu-rep files are too large to travel back)
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_default_final.json 2> gpurun_out/bench_default_final.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference_final.json 2> gpurun_out/bench_reference_final.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file /tmp/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_launches_final.log 2>&1
python tools/launch_summary.py /tmp/launches.csv > gpurun_out/launches_bench_final.txt
cap() { # name, kbench args...
  ncu --set full --clock-control none --import-source on -k regex:tvdk_kernel --launch-skip 3 -c 1 -o /tmp/$1 -f \
      python tools/kbench.py "${@:2}" --reps 3 > gpurun_out/ncu_$1.log 2>&1
  python tools/ncu_summary.py /tmp/$1.ncu-rep --top 25 > gpurun_out/ncu_$1.txt 2>&1
  ncu -i /tmp/$1.ncu-rep --page raw --csv --metrics dram__bytes_read.sum,dram__bytes_write.sum > gpurun_out/ncu_$1_traffic.csv 2>&1
  rm -f /tmp/$1.ncu-rep
}
cap cg_fused_2d --shape 4000 7000 --nb 3 --kernels cg_fused
cap fj1_3d --shape 512 512 512 --nb 1 --kernels fj1_aniso
cap sweepx4_2d --shape 4000 7000 --nb 3 --kernels sweep_x4
cap fj1_hetero_2d --shape 4000 7000 --nb 8 --hetero --kernels fj1_aniso
ls -la gpurun_out | tail -20
python tools/kbench.py --shape 4000 7000 --nb 3 --kernels cg_fused,cg_first,cg_update,cg_pq,cg_init,shrink,fj1_aniso,fj1_iso,fj1x2_aniso,sweep,sweep_x2,sweep_x4 --reps 30 > gpurun_out/kernels_final.txt 2>&1
python tools/kbench.py --shape 512 512 512 --nb 1 --kernels cg_update,cg_pq,cg_init,shrink,fj1_aniso,fj1_iso,sweep --reps 30 >> gpurun_out/kernels_final.txt 2>&1
python tools/kbench.py --shape 512 512 512 --nb 1 --dtype float32 --kernels cg_update,fj1_aniso,fj1_iso --reps 30 >> gpurun_out/kernels_final.txt 2>&1
python tools/kbench.py --shape 4000 7000 --nb 8 --hetero --kernels fj1_aniso,fj1_iso --reps 20 >> gpurun_out/kernels_final.txt 2>&1
cat gpurun_out/kernels_final.txt
