#!/bin/bash
TAG=${1:-r09}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== large sweep"; timeout 600 python tools/large_sweep.py 2>&1 | tee $OUT/large.txt
