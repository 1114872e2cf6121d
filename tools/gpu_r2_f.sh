#!/bin/bash
# Round 2, session F (1 GPU): TMA stencil kernel (cg_pq / sweep 3-D), wave-aware chunks, full test-suite, default bench.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_pytest.log
tail -4 gpurun_out/r2f_pytest.log
for dt in float64 float32; do
  for tma in 1 0; do
    TVDK_TMA=$tma timeout 300 python tools/kbench.py --shape 512 512 512 --nb 1 --dtype $dt --kernels cg_pq,sweep,cg_update,cg_init,shrink --reps 20 2>&1 | grep -E "cg_|sweep|shrink" | sed "s/^/TMA=$tma /"
  done
done > gpurun_out/r2f_stencil_ab.txt 2>&1
cat gpurun_out/r2f_stencil_ab.txt
for w in 1 0; do
  for wl in cfg1_2d_aniso_j1 cfg1_2d_aniso_j20; do
    TVDK_WAVES=$w timeout 300 python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu > gpurun_out/r2f_${wl}_waves${w}.json 2> gpurun_out/r2f_${wl}_waves${w}.err
    python -c "import json,sys; d=json.loads(open('gpurun_out/r2f_${wl}_waves${w}.json').read().strip().splitlines()[-1]); print('$wl waves=$w value', d['value'], 'ms', d['ms_per_step'], 'step frac', d['roofline']['whole_step']['frac'], 'e2e', d['e2e']['value'])"
  done
done
timeout 600 python bench.py --workload cfg4_3d_slab_cg10 --steps 2 --warmup 1 > gpurun_out/r2f_cfg4_n1.json 2> gpurun_out/r2f_cfg4_n1.err; tail -c 900 gpurun_out/r2f_cfg4_n1.json
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2f_bench_default.json 2> gpurun_out/r2f_bench_default.err; echo "bench rc=$?"
python -c "import json; d=json.loads(open('gpurun_out/r2f_bench_default.json').read().strip().splitlines()[-1]); print('default', d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['whole_step']['frac'], {k:(v['value'], v['roofline_per_gpu']['frac']) for k,v in d['slab'].items()})"
