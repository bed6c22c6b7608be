#!/bin/bash
# transform-level barriers incl. named barriers (T = 64 / 128): parity + A/B (libB = CTA-wide barriers)
TAG=${1:-r25}
OUT=gpurun_out/$TAG
mkdir -p $OUT
LIBB=$PWD/chowdsp_fft_b200/lib/ab/libB.so
echo "== pytest gpu" ; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 | tee $OUT/pytest.txt
echo "== sweep A"; timeout 600 python tools/sweep.py --sizes 512,1024,2048,4096 --bytes 2 2>&1 | tee $OUT/sweep_A.txt
echo "== sweep B (cta sync)"; CHOWDSP_FFT_B200_LIB=$LIBB timeout 600 python tools/sweep.py --sizes 512,1024,2048,4096 --bytes 2 2>&1 | tee $OUT/sweep_B.txt
echo "== stft r32"; timeout 600 python bench.py --workload stft --steps 20 --no-e2e --no-cpu 2>&1 | tail -1 | tee $OUT/bench_stft_r32.json
echo "== stft r16"; timeout 600 python bench.py --workload stft --steps 20 --no-e2e --no-cpu --tune radix32_mask=0 2>&1 | tail -1 | tee $OUT/bench_stft_r16.json
echo "== reverb A"; timeout 600 python bench.py --workload reverb --steps 10 --no-e2e --no-cpu 2>&1 | tail -1 | tee $OUT/bench_reverb_A.json
echo "== default"; timeout 600 python bench.py --no-e2e --no-cpu 2>&1 | tail -1 | tee $OUT/bench_default.json
echo "== unordered"; timeout 600 python bench.py --workload c2c4096_unordered --no-e2e --no-cpu 2>&1 | tail -1 | tee $OUT/bench_c2c4096_unordered.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fft_kernel -s 3 -c 1 -f -o $OUT/prof_stft \
   python bench.py --workload stft --steps 2 --warmup 3 --no-e2e --no-cpu > $OUT/prof_stft.log 2>&1
ls -la $OUT
