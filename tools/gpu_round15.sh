#!/bin/bash
TAG=${1:-r15}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest radix32"; timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "radix32 or large" 2>&1 | tail -8 | tee $OUT/pytest.txt
echo "== sweep r16"; timeout 600 python tools/sweep.py --sizes 512,1024,2048,8192,16384,32768 --bytes 2 --radix32-mask 0 2>&1 | tee $OUT/sweep_r16.txt
echo "== sweep r32"; timeout 600 python tools/sweep.py --sizes 512,1024,2048,8192,16384,32768 --bytes 2 --radix32-mask 32767 2>&1 | tee $OUT/sweep_r32.txt
echo "== bench stft r32"; timeout 600 python bench.py --workload stft --steps 10 --no-e2e --no-cpu --tune radix32_mask=0x7fff 2>&1 | tail -1 | tee $OUT/bench_stft_r32.json
echo "== large jfast8"; timeout 600 python tools/large_sweep.py --tile-c-jfast=8 15 16 18 24 26 2>&1 | tee $OUT/large_jfast8.txt
echo "== large jfast16"; timeout 600 python tools/large_sweep.py --tile-c-jfast=16 15 16 18 24 26 2>&1 | tee $OUT/large_jfast16.txt
prof() { # name workload regex skip count tune
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$3 -s $4 -c $5 -f -o $OUT/prof_$1 \
     python bench.py --workload $2 --steps 1 --warmup 3 --no-e2e --no-cpu --tune radix32_mask=0x7fff > $OUT/prof_$1.log 2>&1
}
prof stft_r32 stft stft_kernel\|fft_kernel 3 1
prof c2c16384_r32 c2c16384 fft_kernel 3 1
prof c2c8192_r32 c2c8192 fft_kernel 3 1
ls -la $OUT
