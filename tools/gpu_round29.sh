#!/bin/bash
# mixed-radix sizes: parity + informational throughput; config 1 latency
TAG=${1:-r29}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest gpu" ; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 | tee $OUT/pytest.txt
echo "== sweep mixed"; timeout 600 python tools/sweep.py --sizes 96,768,1920,9216 --bytes 1 2>&1 | tee $OUT/sweep_mixed.txt
echo "== single1024"; timeout 600 python bench.py --workload single1024 2>&1 | tail -1 | tee $OUT/bench_single1024.json
ls -la $OUT
