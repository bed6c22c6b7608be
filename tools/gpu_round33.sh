#!/bin/bash
# unordered complex spectra stored / loaded directly (pair shuffle, no staging image): parity + A/B (libB = staging image)
TAG=${1:-r33}
OUT=gpurun_out/$TAG
mkdir -p $OUT
LIBB=$PWD/chowdsp_fft_b200/lib/libB_staged.so
echo "== pytest" ; timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee $OUT/pytest_gpu.txt
echo "== sweep A"; timeout 600 python tools/sweep.py --sizes 128,256,512,1024,2048,4096,8192 --kinds c --bytes 2 2>&1 | grep -E "unordered" | tee $OUT/sweep_A.txt
echo "== sweep B"; CHOWDSP_FFT_B200_LIB=$LIBB timeout 600 python tools/sweep.py --sizes 128,256,512,1024,2048,4096,8192 --kinds c --bytes 2 2>&1 | grep -E "unordered" | tee $OUT/sweep_B.txt
echo "== bench unordered A"; timeout 300 python bench.py --workload c2c4096_unordered --no-e2e --no-cpu 2>&1 | tail -1 | tee $OUT/bench_c2c4096_unordered_A.json
echo "== bench unordered B"; CHOWDSP_FFT_B200_LIB=$LIBB timeout 300 python bench.py --workload c2c4096_unordered --no-e2e --no-cpu 2>&1 | tail -1 | tee $OUT/bench_c2c4096_unordered_B.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fft_kernel -s 3 -c 1 -f -o $OUT/prof_c2c4096_unordered \
   python bench.py --workload c2c4096_unordered --steps 2 --warmup 3 --no-e2e --no-cpu > $OUT/prof.log 2>&1
ls -la $OUT
