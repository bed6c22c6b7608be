#!/bin/bash
# L2 prefetch-ahead for the tile passes of the large transforms (tuning hook tile_pf = distance in tiles)
TAG=${1:-r38}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for pf in 0 148 296 444 592 1184; do
  echo "== huge tile_pf=$pf"; timeout 300 python bench.py --workload huge --no-e2e --no-cpu --tune tile_pf=$pf 2>&1 | tail -1 | tee $OUT/bench_huge_pf$pf.json
done
for pf in 0 296 592; do
  echo "== sweep tile_pf=$pf"; CFB_TUNE=tile_pf=$pf timeout 600 python tools/large_sweep.py 16 20 24 26 2>&1 | tee $OUT/large_pf$pf.txt
done
echo "== pytest large (pf 296)"; CFB_TUNE_PYTEST=1 timeout 600 python - <<'PY' 2>&1 | tail -3 | tee $OUT/pytest_large_pf.txt
import sys, pytest
import chowdsp_fft_b200 as cf
cf.set_tuning("tile_pf", 296)
sys.exit(pytest.main(["tests/test_gpu_parity.py", "-x", "-q", "-m", "gpu", "-k", "large or config5"]))
PY
