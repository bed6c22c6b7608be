#!/bin/bash
OUT=gpurun_out/r2m
mkdir -p $OUT
echo "== pytest mixed"; timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "mixed or restated or every_size" 2>&1 | tail -4 | tee $OUT/pytest_mixed.txt
SZ=96,160,192,288,384,480,640,768,1920,2560,9216,12288
echo "== sweep mixq real"; timeout 600 python tools/sweep.py --sizes $SZ --kinds r --bytes 2 --pause 0.5 --repeats 2 --layouts ordered,w8 2>&1 | grep -E "C2C|R2C|C2R" | tee $OUT/sweep_mixq_real.txt
for v in A B; do
  lib=chowdsp_fft_b200/lib/ab/lib$v.so; [ $v = A ] && lib=chowdsp_fft_b200/lib/libchowdsp_fft_b200.so
  echo "== small sizes, lib $v"; CHOWDSP_FFT_B200_LIB=$PWD/$lib timeout 600 python tools/sweep.py --bytes 2 --steps 20 --pause 0.5 --repeats 2 --sizes 32,64,128 --layouts ordered,w4 2>&1 | grep -E "C2C|R2C|C2R" | tee $OUT/sweep_small_$v.txt
done
