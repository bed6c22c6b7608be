#!/bin/bash
# mixed-radix kernel at 64 registers (2 CTAs of 512 threads per SM): parity + sweep
TAG=${1:-r42}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest mixed" ; timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "mixed or reference_test_suite" 2>&1 | tail -4 | tee $OUT/pytest.txt
echo "== sweep"; timeout 600 python tools/sweep.py --sizes 96,192,384,480,640,768,1920,9216,12288 --bytes 2 2>&1 | grep -E "C2C|R2C|C2R" | tee $OUT/sweep_mixed.txt
