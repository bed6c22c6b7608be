#!/bin/bash
# Round-2 diagnostic: does the consolidated sweep run power-capped?  (clock / power per cell, with and without a pause
# between cells) + ncu captures of the real-transform kernels that sit under 5.6 TB/s.
OUT=gpurun_out/r2g
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit,temperature.gpu --format=csv > $OUT/gpu.txt 2>&1
echo "== sweep back to back"; timeout 600 python tools/sweep.py --bytes 2 --steps 20 --sizes 2048,4096,8192,16384,32768 --layouts ordered,w8 2>&1 | grep -E "C2C|R2C|C2R" | tee $OUT/sweep_b2b.txt
echo "== sweep with 1 s pause, best of 3 groups"; timeout 900 python tools/sweep.py --bytes 2 --steps 20 --pause 1.0 --repeats 3 --sizes 2048,4096,8192,16384,32768 --layouts ordered,w8 2>&1 | grep -E "C2C|R2C|C2R" | tee $OUT/sweep_pause.txt
cap() { # name, size, kind, layout, skip
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"fft_kernel|pipe_kernel" -s $5 -c 1 -f -o $OUT/prof_$1 \
     python tools/sweep.py --bytes 2 --steps 2 --sizes $2 --kinds $3 --layouts $4 > $OUT/prof_$1.log 2>&1
}
echo "== ncu"
cap r2c4096 4096 r ordered 3
cap c2r4096 4096 r ordered 8
cap c2r8192 8192 r ordered 8
cap r2c8192_w8 8192 r w8 3
cap r2c32768 32768 r ordered 3
cap c2c16384 16384 c ordered 3
ls -la $OUT
