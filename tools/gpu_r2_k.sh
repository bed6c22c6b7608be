#!/bin/bash
# Q x 2^p mixed-radix kernel: parity through the C ABI + sweep against the generic kernel
OUT=gpurun_out/r2k
mkdir -p $OUT
echo "== pytest mixed"; timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "mixed or juce or restated" 2>&1 | tail -5 | tee $OUT/pytest_mixed.txt
SZ=96,160,192,288,384,480,640,768,1920,2560,9216,12288
echo "== sweep mixq"; timeout 600 python tools/sweep.py --sizes $SZ --bytes 2 --pause 0.5 --repeats 2 --layouts ordered,w8 2>&1 | grep -E "C2C|R2C|C2R" | tee $OUT/sweep_mixq.txt
echo "== sweep generic"; timeout 600 python tools/sweep.py --sizes $SZ --bytes 2 --pause 0.5 --repeats 2 --layouts ordered,w8 --tune mixq=0 2>&1 | grep -E "C2C|R2C|C2R" | tee $OUT/sweep_generic.txt
