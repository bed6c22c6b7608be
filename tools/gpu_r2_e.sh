#!/bin/bash
# round-2 call E: tile passes with 32 points per thread (tile_r), ncu of the three passes of 2^28
OUT=gpurun_out/r2e
mkdir -p $OUT
echo "== parity tile_r"; timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "32_points_per_thread or large_transforms" 2>&1 | tail -6 | tee $OUT/pytest_tile_r.txt
echo "== sweep tile_r=0"; timeout 600 python tools/large_sweep.py 16 18 20 22 24 26 28 2>&1 | tee $OUT/sweep_tile_r0.txt
echo "== sweep tile_r=1"; CFB_TUNE=tile_r=1 timeout 600 python tools/large_sweep.py 16 18 20 22 24 26 28 2>&1 | tee $OUT/sweep_tile_r1.txt
echo "== sweep tile_r=1 tile_c=8"; CFB_TUNE=tile_r=1,tile_c=8 timeout 600 python tools/large_sweep.py --complex-only 20 24 28 2>&1 | tee $OUT/sweep_tile_r1_c8.txt
echo "== ncu huge r0"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tile_fft_kernel -s 6 -c 3 -f -o $OUT/prof_huge_r0 \
   python tools/large_sweep.py --complex-only 28 > $OUT/prof_huge_r0.log 2>&1
echo "== ncu huge r1"
CFB_TUNE=tile_r=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:tile_fft_kernel -s 6 -c 3 -f -o $OUT/prof_huge_r1 \
   python tools/large_sweep.py --complex-only 28 > $OUT/prof_huge_r1.log 2>&1
ls -la $OUT
