#!/bin/bash
# Round 2, session D (1 GPU): alternating chunk direction A/B on the small-image configs, streamed series at N = 1.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_pytest.log
tail -4 gpurun_out/r2d_pytest.log
for alt in 1 0; do
  for wl in cfg1_2d_aniso_j1 cfg1_2d_aniso_j20; do
    TVDK_ALTERNATE=$alt timeout 300 python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu > gpurun_out/r2d_${wl}_alt${alt}.json 2> gpurun_out/r2d_${wl}_alt${alt}.err
    python -c "import json,sys; d=json.loads(open('gpurun_out/r2d_${wl}_alt${alt}.json').read().strip().splitlines()[-1]); print('$wl alt=$alt value', d['value'], 'ms', d['ms_per_step'], 'step frac', d['roofline']['whole_step']['frac'], 'e2e', d['e2e']['value'])"
  done
done
TVDK_ALTERNATE=0 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-slab > gpurun_out/r2d_cfg2_alt0.json 2>&1
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-slab > gpurun_out/r2d_cfg2_alt1.json 2>&1
for a in 0 1; do python -c "import json; d=json.loads(open('gpurun_out/r2d_cfg2_alt$a.json').read().strip().splitlines()[-1]); print('cfg2 alt=$a', d['value'], d['ms_per_step'])"; done
timeout 900 python bench.py --workload cfg3_series512 --steps 1 --warmup 1 > gpurun_out/r2d_cfg3_series512_n1.json 2> gpurun_out/r2d_cfg3_series512_n1.err; echo "rc=$?"
tail -c 2500 gpurun_out/r2d_cfg3_series512_n1.json; tail -5 gpurun_out/r2d_cfg3_series512_n1.err
