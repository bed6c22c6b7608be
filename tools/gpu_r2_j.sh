#!/bin/bash
# tensor-map L2 prefetch for the tile passes (tile_pf) + GPU parity suite after the derived-twiddle change
OUT=gpurun_out/r2j
mkdir -p $OUT
for pf in 0 148 296 444 592 888; do
  echo "== large sweep tile_pf=$pf"; CFB_TUNE=tile_pf=$pf timeout 300 python tools/large_sweep.py --complex-only 16 18 20 22 24 26 28 2>&1 | grep -E "C2C" | tee $OUT/sweep_large_pf$pf.txt
done
echo "== bench huge pf 296 (parity inside)"; timeout 300 python bench.py --workload huge --no-cpu --no-e2e --tune tile_pf=296 2>&1 | tail -1 | tee $OUT/bench_huge_pf296.json
echo "== pytest -m gpu"; ( time timeout 1500 python -m pytest tests -x -q -m gpu ) 2>&1 | tail -8 | tee $OUT/pytest_gpu.txt
