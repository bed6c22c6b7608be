#!/bin/bash
# wpipe / wistft with the landing buffer doubling as the exchange buffer (A: 16 / 12 warps per SM) vs separate buffers (libB: 13 / 10)
TAG=${1:-r35}
OUT=gpurun_out/$TAG
mkdir -p $OUT
LIBB=$PWD/chowdsp_fft_b200/lib/libB_noalias.so
echo "== pytest A" ; timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "warp_pipelined or stft or istft" 2>&1 | tail -4 | tee $OUT/pytest_A.txt
for rep in 1 2; do
echo "== stft A"; timeout 300 python bench.py --workload stft --no-e2e --no-cpu 2>&1 | tail -1 | tee $OUT/bench_stft_A$rep.json
echo "== stft B"; CHOWDSP_FFT_B200_LIB=$LIBB timeout 300 python bench.py --workload stft --no-e2e --no-cpu 2>&1 | tail -1 | tee $OUT/bench_stft_B$rep.json
echo "== istft A"; timeout 300 python bench.py --workload istft --no-e2e --no-cpu 2>&1 | tail -1 | tee $OUT/bench_istft_A$rep.json
echo "== istft B"; CHOWDSP_FFT_B200_LIB=$LIBB timeout 300 python bench.py --workload istft --no-e2e --no-cpu 2>&1 | tail -1 | tee $OUT/bench_istft_B$rep.json
done
for w in 14 12; do
echo "== stft A warps=$w"; timeout 300 python bench.py --workload stft --no-e2e --no-cpu --tune wpipe=$((2 + w*256)) 2>&1 | tail -1 | tee $OUT/bench_stft_A_w$w.json
done
echo "== istft A warps=10"; timeout 300 python bench.py --workload istft --no-e2e --no-cpu --tune wistft=$((1 + 10*256)) 2>&1 | tail -1 | tee $OUT/bench_istft_A_w10.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wpipe_kernel -s 3 -c 1 -f -o $OUT/prof_stft_wpipe \
   python bench.py --workload stft --steps 2 --warmup 3 --no-e2e --no-cpu > $OUT/prof_stft_wpipe.log 2>&1
ls -la $OUT
