#!/bin/bash
TAG=${1:-r03}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest -m gpu" ; timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee $OUT/pytest_gpu.txt
echo "== sweep" ; timeout 900 python tools/sweep.py --bytes 2 --sizes 256,1024,2048,4096,8192,16384,32768 --json $OUT/sweep.json 2>&1 | tee $OUT/sweep.txt
ls -la $OUT
