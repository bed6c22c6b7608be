#!/bin/bash
TAG=${1:-r04}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest -m gpu" ; timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee $OUT/pytest_gpu.txt
for pf in 0 1; do
echo "== sweep prefetch=$pf" ; CHOWDSP_FFT_B200_L2_PREFETCH=$pf timeout 900 python tools/sweep.py --bytes 2 --sizes 1024,4096,8192,16384,32768 --json $OUT/sweep_pf$pf.json 2>&1 | grep -E "fwd" | tee $OUT/sweep_pf$pf.txt
done
echo "== bench (default)" ; timeout 900 python bench.py --no-cpu --no-e2e 2>&1 | tail -1 | cut -c1-400 | tee $OUT/bench_default.json
