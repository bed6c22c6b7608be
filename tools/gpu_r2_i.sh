#!/bin/bash
# burst-mode re-tune of the dispatch hooks + A/B of two compile-time switches (derived real twiddles, direct unordered-real access)
OUT=gpurun_out/r2i
mkdir -p $OUT
echo "== retune"; timeout 900 python tools/retune.py 2>&1 | tee $OUT/retune.txt
for v in A B C; do
  lib=chowdsp_fft_b200/lib/ab/lib$v.so; [ $v = A ] && lib=chowdsp_fft_b200/lib/libchowdsp_fft_b200.so
  echo "== real kinds, lib $v"; CHOWDSP_FFT_B200_LIB=$PWD/$lib timeout 600 python tools/sweep.py --bytes 2 --steps 20 --pause 1.0 --repeats 3 --kinds r --sizes 512,1024,2048,4096,8192,16384,32768 --layouts ordered,w8 2>&1 | grep -E "R2C|C2R" | tee $OUT/sweep_real_$v.txt
done
