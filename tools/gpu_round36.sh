#!/bin/bash
# strided tile-copy ceiling (memory pattern of the large-transform passes); full GPU suite on the fenced alias build; stft / istft lines
TAG=${1:-r36}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== strided copy probe"; timeout 120 tools/probes/strided_copy_probe 2>&1 | tee $OUT/strided_copy.txt
echo "== pytest -m gpu (all)" ; timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee $OUT/pytest_gpu.txt
echo "== stft"; timeout 300 python bench.py --workload stft --no-e2e --no-cpu 2>&1 | tail -1 | tee $OUT/bench_stft.json
echo "== istft"; timeout 300 python bench.py --workload istft --no-e2e --no-cpu 2>&1 | tail -1 | tee $OUT/bench_istft.json
