#!/bin/bash
# shuffle-based mirror exchange of the real split / merge step (T <= 32): full GPU parity suite + A/B (libB = shared-memory exchange)
TAG=${1:-r29}
OUT=gpurun_out/$TAG
mkdir -p $OUT
LIBB=$PWD/chowdsp_fft_b200/lib/libB_noshfl.so
echo "== pytest -m gpu" ; timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 | tee $OUT/pytest_gpu.txt
echo "== stft A"; timeout 300 python bench.py --workload stft --no-e2e --no-cpu 2>&1 | tail -1 | tee $OUT/bench_stft_A.json
echo "== stft B"; CHOWDSP_FFT_B200_LIB=$LIBB timeout 300 python bench.py --workload stft --no-e2e --no-cpu 2>&1 | tail -1 | tee $OUT/bench_stft_B.json
echo "== istft A"; timeout 300 python bench.py --workload istft --no-e2e --no-cpu 2>&1 | tail -1 | tee $OUT/bench_istft_A.json
echo "== istft B"; CHOWDSP_FFT_B200_LIB=$LIBB timeout 300 python bench.py --workload istft --no-e2e --no-cpu 2>&1 | tail -1 | tee $OUT/bench_istft_B.json
echo "== sweep A"; timeout 600 python tools/sweep.py --sizes 64,128,256,512,1024,2048 --kinds r --bytes 2 2>&1 | grep -E "R2C|C2R" | tee $OUT/sweep_A.txt
echo "== sweep B"; CHOWDSP_FFT_B200_LIB=$LIBB timeout 600 python tools/sweep.py --sizes 64,128,256,512,1024,2048 --kinds r --bytes 2 2>&1 | grep -E "R2C|C2R" | tee $OUT/sweep_B.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wpipe_kernel -s 3 -c 1 -f -o $OUT/prof_stft_wpipe \
   python bench.py --workload stft --steps 2 --warmup 3 --no-e2e --no-cpu > $OUT/prof_stft_wpipe.log 2>&1
ls -la $OUT
