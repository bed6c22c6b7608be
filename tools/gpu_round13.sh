#!/bin/bash
# Tile kernels v2 (smem twiddle rows, tile width 8/16): parity + sweep of both widths + huge bench + ncu.
TAG=${1:-r13}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest large/dist"; timeout 900 python -m pytest tests -x -q -m gpu -k "large or config5 or dist" 2>&1 | tail -8 | tee $OUT/pytest_large.txt
echo "== large sweep tile_c=8"; timeout 600 python tools/large_sweep.py --tile-c=8 15 16 18 20 24 28 2>&1 | tee $OUT/large_c8.txt
echo "== large sweep tile_c=16"; timeout 600 python tools/large_sweep.py --tile-c=16 15 16 18 20 24 28 2>&1 | tee $OUT/large_c16.txt
echo "== bench huge"; timeout 600 python bench.py --workload huge --steps 10 --no-e2e --no-cpu 2>&1 | tail -1 | tee $OUT/bench_huge.json
echo "== bench single1024"; timeout 600 python bench.py --workload single1024 --steps 200 2>&1 | tail -1 | tee $OUT/bench_single1024.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches_huge.csv \
   python bench.py --workload huge --steps 2 --warmup 3 --no-e2e --no-cpu > $OUT/launches_huge.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tile_fft -s 9 -c 3 -f -o $OUT/prof_huge \
     python bench.py --workload huge --steps 1 --warmup 3 --no-e2e --no-cpu > $OUT/prof_huge.log 2>&1
ls -la $OUT
