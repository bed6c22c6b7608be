#!/bin/bash
# pipe kernel v2 (smem twiddle rows, derived real twiddles, 2 CTAs/SM at 2^13) + real_tw A/B (libB = table loads)
TAG=${1:-r19}
OUT=gpurun_out/$TAG
mkdir -p $OUT
LIBB=$PWD/chowdsp_fft_b200/lib/ab/libB.so
echo "== pytest parity" ; timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -8 | tee $OUT/pytest.txt
echo "== sweep A (real_tw derived)"; timeout 600 python tools/sweep.py --sizes 512,1024,2048,4096,8192,16384,32768 --bytes 2 2>&1 | tee $OUT/sweep_A.txt
echo "== sweep B (table)"; CHOWDSP_FFT_B200_LIB=$LIBB timeout 600 python tools/sweep.py --sizes 512,1024,2048,4096,8192,16384,32768 --bytes 2 --kinds r 2>&1 | tee $OUT/sweep_B.txt
echo "== stft A"; timeout 600 python bench.py --workload stft --steps 10 --no-e2e --no-cpu 2>&1 | tail -1 | tee $OUT/bench_stft_A.json
echo "== stft B"; CHOWDSP_FFT_B200_LIB=$LIBB timeout 600 python bench.py --workload stft --steps 10 --no-e2e --no-cpu 2>&1 | tail -1 | tee $OUT/bench_stft_B.json
echo "== reverb A"; timeout 600 python bench.py --workload reverb --steps 10 --no-e2e --no-cpu 2>&1 | tail -1 | tee $OUT/bench_reverb_A.json
echo "== reverb B"; CHOWDSP_FFT_B200_LIB=$LIBB timeout 600 python bench.py --workload reverb --steps 10 --no-e2e --no-cpu 2>&1 | tail -1 | tee $OUT/bench_reverb_B.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pipe_kernel -s 3 -c 1 -f -o $OUT/prof_c2c16384_pipe \
   python bench.py --workload c2c16384 --steps 2 --warmup 3 --no-e2e --no-cpu > $OUT/prof_c2c16384_pipe.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pipe_kernel -s 3 -c 1 -f -o $OUT/prof_c2c8192_pipe \
   python bench.py --workload c2c8192 --steps 2 --warmup 3 --no-e2e --no-cpu > $OUT/prof_c2c8192_pipe.log 2>&1
ls -la $OUT
