#!/bin/bash
TAG=${1:-r02}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest -m gpu" ; timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee $OUT/pytest_gpu.txt
echo "== sweep" ; timeout 900 python tools/sweep.py --json $OUT/sweep.json 2>&1 | tee $OUT/sweep.txt
echo "== bench (default)" ; timeout 900 python bench.py 2>&1 | tail -1 | tee $OUT/bench_default.json
echo "== ncu full (c2c4096 ordered + unordered, c2c16384, r2c8192 unordered)"
for wl in c2c4096 c2c4096_unordered c2c16384 r2c8192; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fft_kernel -s 3 -c 1 -f -o $OUT/prof_$wl \
   python bench.py --workload $wl --steps 2 --warmup 3 --no-e2e --no-cpu > $OUT/prof_$wl.log 2>&1
done
ls -la $OUT
