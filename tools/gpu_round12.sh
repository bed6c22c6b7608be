#!/bin/bash
# STFT gather kernel validation + ncu captures of the kernels below 0.8 of the roofline.
TAG=${1:-r12}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest stft"; timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "stft or strided" 2>&1 | tail -8 | tee $OUT/pytest_stft.txt
echo "== bench stft"; timeout 600 python bench.py --workload stft --steps 10 --no-e2e --no-cpu 2>&1 | tail -1 | tee $OUT/bench_stft.json
prof() { # name workload regex skip count
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$3 -s $4 -c $5 -f -o $OUT/prof_$1 \
     python bench.py --workload $2 --steps 1 --warmup 3 --no-e2e --no-cpu > $OUT/prof_$1.log 2>&1
}
prof huge huge tile_fft 9 3
prof c2c8192 c2c8192 fft_kernel 3 1
prof c2c16384 c2c16384 fft_kernel 3 1
prof stft stft stft_kernel 3 1
prof c2c4096u c2c4096_unordered fft_kernel 3 1
prof c2c4096 c2c4096 fft_kernel 3 1
ls -la $OUT
