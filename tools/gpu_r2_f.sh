#!/bin/bash
# round-2 call F: persistent tensor-map TMA tile kernel: parity, A/B sweep, ncu
OUT=gpurun_out/r2f
mkdir -p $OUT
echo "== parity"; timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tma_tile" 2>&1 | tail -8 | tee $OUT/pytest_tma.txt
echo "== sweep tile_tma=0"; timeout 600 python tools/large_sweep.py 16 18 20 22 24 26 28 2>&1 | tee $OUT/sweep_tma0.txt
echo "== sweep tile_tma=1"; CFB_TUNE=tile_tma=1 timeout 600 python tools/large_sweep.py 16 18 20 22 24 26 28 2>&1 | tee $OUT/sweep_tma1.txt
echo "== sweep tile_tma=1 no chunk"; CFB_TUNE=tile_tma=1,l2_chunk_mb=0 timeout 600 python tools/large_sweep.py --complex-only 16 20 24 2>&1 | tee $OUT/sweep_tma1_nochunk.txt
echo "== ncu huge tma"
CFB_TUNE=tile_tma=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:tile_tma_kernel -s 6 -c 3 -f -o $OUT/prof_huge_tma \
   python tools/large_sweep.py --complex-only 28 > $OUT/prof_huge_tma.log 2>&1
ls -la $OUT
