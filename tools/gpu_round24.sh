#!/bin/bash
# warp-level transform barriers (T <= 32): parity + A/B (libB = CTA-wide barriers), STFT pipe A/B again
TAG=${1:-r24}
OUT=gpurun_out/$TAG
mkdir -p $OUT
LIBB=$PWD/chowdsp_fft_b200/lib/ab/libB.so
echo "== pytest gpu" ; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 | tee $OUT/pytest.txt
echo "== sweep A (warp sync)"; timeout 600 python tools/sweep.py --sizes 64,128,256,512,1024,2048 --bytes 2 2>&1 | tee $OUT/sweep_A.txt
echo "== sweep B (cta sync)"; CHOWDSP_FFT_B200_LIB=$LIBB timeout 600 python tools/sweep.py --sizes 64,128,256,512,1024,2048 --bytes 2 2>&1 | tee $OUT/sweep_B.txt
for v in A B; do
  if [ $v = B ]; then export CHOWDSP_FFT_B200_LIB=$LIBB; fi
  echo "== stft pipe $v"; timeout 600 python bench.py --workload stft --steps 20 --no-e2e --no-cpu 2>&1 | tail -1 | tee $OUT/bench_stft_pipe_$v.json
  echo "== stft nopipe $v"; timeout 600 python bench.py --workload stft --steps 20 --no-e2e --no-cpu --tune stft_pipe=0 2>&1 | tail -1 | tee $OUT/bench_stft_nopipe_$v.json
done
unset CHOWDSP_FFT_B200_LIB
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stft_pipe_kernel -s 3 -c 1 -f -o $OUT/prof_stft_pipe \
   python bench.py --workload stft --steps 2 --warmup 3 --no-e2e --no-cpu > $OUT/prof_stft_pipe.log 2>&1
ls -la $OUT
