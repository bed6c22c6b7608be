#!/bin/bash
# mixed-radix kernel: radix-16 passes, 512-thread cap, first / last pass fused with the global load / store: full parity suite + sweep
TAG=${1:-r41}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest -m gpu (all)" ; timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee $OUT/pytest_gpu.txt
echo "== sweep"; timeout 600 python tools/sweep.py --sizes 96,192,384,480,640,768,1920,9216,12288 --bytes 2 2>&1 | grep -E "C2C|R2C|C2R" | tee $OUT/sweep_mixed.txt
