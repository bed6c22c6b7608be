#!/bin/bash
# after the derived-twiddle policy + mixq: full GPU suite, burst-mode re-tune, mixed sweep
OUT=gpurun_out/r2l
mkdir -p $OUT
echo "== pytest -m gpu"; ( time timeout 1500 python -m pytest tests -x -q -m gpu ) 2>&1 | tail -8 | tee $OUT/pytest_gpu.txt
echo "== retune"; timeout 900 python tools/retune.py 2>&1 | tee $OUT/retune.txt
SZ=96,160,192,288,384,480,640,768,1920,2560,9216,12288
echo "== sweep mixq"; timeout 600 python tools/sweep.py --sizes $SZ --bytes 2 --pause 0.5 --repeats 2 --layouts ordered,w8 2>&1 | grep -E "C2C|R2C|C2R" | tee $OUT/sweep_mixq.txt
