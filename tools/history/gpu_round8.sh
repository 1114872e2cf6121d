#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -k "golden or seams" > gpurun_out/pytest_parity_v8.log 2>&1; tail -3 gpurun_out/pytest_parity_v8.log
out=gpurun_out/x2.txt; : > $out
for lib in "" "$PWD/darsia_b200/csrc/libtvdk_minb3.so" "$PWD/darsia_b200/csrc/libtvdk_minb4.so"; do
echo "== lib=${lib:-default}" >> $out
TVDK_LIB=$lib python tools/kbench.py --shape 4000 7000 --nb 3 --kernels fj1_aniso,fj1x2_aniso,fj1_iso,fj1x2_iso --reps 30 >> $out 2>&1
TVDK_LIB=$lib python tools/kbench.py --shape 4000 7000 --nb 8 --hetero --kernels fj1_aniso,fj1x2_aniso --reps 20 >> $out 2>&1
TVDK_LIB=$lib python tools/kbench.py --shape 1788 3180 --nb 1 --kernels fj1_aniso,fj1x2_aniso --reps 100 >> $out 2>&1
TVDK_LIB=$lib python tools/kbench.py --shape 4000 7000 --nb 3 --dtype float32 --kernels fj1_aniso,fj1x2_aniso --reps 30 >> $out 2>&1
done
echo "== tiles (default lib)" >> $out
python tools/kbench.py --shape 4000 7000 --nb 3 --kernels fj1x2_aniso --tile-x 64,128 --zchunk 32,64,128 --reps 30 >> $out 2>&1
cat $out
