#!/bin/bash
# Re-entry check of HEAD: GPU parity suite, smoke, default bench (with e2e + cpu), reference arm, STFT line.
TAG=${1:-r26}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nproc > $OUT/nproc.txt
echo "== pytest -m gpu" ; timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 | tee $OUT/pytest_gpu.txt
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/smoke.txt
echo "== bench (default)" ; timeout 600 python bench.py 2>&1 | tail -1 | tee $OUT/bench_default.json
echo "== bench reference arm" ; timeout 600 python bench.py --impl reference --steps 5 --warmup 1 2>&1 | tail -1 | tee $OUT/bench_reference.json
echo "== stft"; timeout 300 python bench.py --workload stft --no-e2e --no-cpu 2>&1 | tail -1 | tee $OUT/bench_stft.json
