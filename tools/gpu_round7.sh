#!/bin/bash
TAG=${1:-r08}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest -m gpu (large only)" ; timeout 1500 python -m pytest tests -x -q -m gpu -k "large or config5 or restated" 2>&1 | tail -5 | tee $OUT/pytest_gpu.txt
echo "== large sweep"; timeout 600 python tools/large_sweep.py 2>&1 | tee $OUT/large.txt
echo "== ncu tile kernels"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tile_fft -s 6 -c 3 -f -o $OUT/prof_tile24 python tools/large_sweep.py 24 > $OUT/prof_tile24.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tile_fft -s 4 -c 2 -f -o $OUT/prof_tile20 python tools/large_sweep.py 20 > $OUT/prof_tile20.log 2>&1
ls -la $OUT
