#!/bin/bash
# Round 2, session I (1 GPU): programmatic dependent launch A/B (TVDK_PDL=0/1) -- the GPU test-suite with it ON, then
# cfg1 (75 us launches), cfg2 (headline), and the single-GPU slab loops of cfg4 / cfg5.
set -x
mkdir -p gpurun_out
TVDK_PDL=1 timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2i_pytest_pdl1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2i_pytest_pdl1.log
tail -4 gpurun_out/r2i_pytest_pdl1.log
for wl in cfg1_2d_aniso_j1 cfg1_2d_aniso_j20 cfg2_2d_iso_cg10; do
  for pdl in 0 1 0 1; do
    TVDK_PDL=$pdl timeout 600 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu --no-slab > gpurun_out/r2i_${wl}_pdl${pdl}.json 2> gpurun_out/r2i_${wl}_pdl${pdl}.err
    python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r2i_${wl}_pdl${pdl}.json") if l.startswith("{")][-1])
print("$wl pdl=$pdl value", round(d["value"],1), "ms_per_step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), "whole_step_frac", round(d["roofline"]["whole_step"]["frac"],4), "dom_frac", round(d["roofline"]["frac"],4))
PY
  done
done
for wl in cfg4_3d_slab_cg10 cfg5_3d_slab_j1_f32; do
  for pdl in 0 1; do
    TVDK_PDL=$pdl timeout 600 python bench.py --workload $wl --steps 2 --warmup 1 --no-cpu > gpurun_out/r2i_${wl}_pdl${pdl}.json 2> gpurun_out/r2i_${wl}_pdl${pdl}.err
    python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r2i_${wl}_pdl${pdl}.json") if l.startswith("{")][-1])
print("$wl pdl=$pdl value", round(d["value"],1), "ms_per_step", round(d["ms_per_step"],3), "frac", round(d["roofline"]["frac"],4))
PY
  done
done
