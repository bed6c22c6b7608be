#!/bin/bash
# tile_pipe_kernel (persistent TMA-staged tile passes of the large transforms): parity + A/B; wistft with 12 warps.
TAG=${1:-r31}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest large" ; timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "large or config5 or istft" 2>&1 | tail -6 | tee $OUT/pytest_large.txt
echo "== huge pipe (default tile widths)"; timeout 300 python bench.py --workload huge --no-e2e --no-cpu 2>&1 | tail -1 | tee $OUT/bench_huge_pipe.json
echo "== huge pipe tile_c=8"; timeout 300 python bench.py --workload huge --no-e2e --no-cpu --tune tile_c=8 2>&1 | tail -1 | tee $OUT/bench_huge_pipe_c8.json
echo "== huge no pipe"; timeout 300 python bench.py --workload huge --no-e2e --no-cpu --tune tile_pipe=0 2>&1 | tail -1 | tee $OUT/bench_huge_nopipe.json
echo "== sweep pipe"; timeout 600 python tools/large_sweep.py 15 16 18 20 22 24 26 2>&1 | tee $OUT/large_pipe.txt
echo "== sweep pipe c8"; CFB_TUNE=tile_c=8 timeout 600 python tools/large_sweep.py 15 16 18 20 22 24 26 2>&1 | tee $OUT/large_pipe_c8.txt
echo "== sweep no pipe"; CFB_TUNE=tile_pipe=0 timeout 600 python tools/large_sweep.py 15 16 18 20 22 24 26 2>&1 | tee $OUT/large_nopipe.txt
echo "== istft 12 warps"; timeout 300 python bench.py --workload istft --no-e2e --no-cpu 2>&1 | tail -1 | tee $OUT/bench_istft.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tile_pipe_kernel -s 6 -c 3 -f -o $OUT/prof_huge_tile_pipe \
   python bench.py --workload huge --steps 2 --warmup 3 --no-e2e --no-cpu > $OUT/prof_huge.log 2>&1
ls -la $OUT
