#!/bin/bash
# unordered REAL spectra: 4-byte direct global accesses (libB, -DCFB_UNORD_REAL_DIRECT=1) vs staging image (A): parity of B + A/B sweep
TAG=${1:-r34}
OUT=gpurun_out/$TAG
mkdir -p $OUT
LIBB=$PWD/chowdsp_fft_b200/lib/libB_rdirect.so
echo "== pytest B" ; CHOWDSP_FFT_B200_LIB=$LIBB timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "every_size or golden or convolve or impulses or batched_equals" 2>&1 | tail -4 | tee $OUT/pytest_B.txt
echo "== sweep A"; timeout 600 python tools/sweep.py --sizes 256,512,1024,2048,4096,8192,16384 --kinds r --bytes 2 2>&1 | grep -E "unordered" | tee $OUT/sweep_A.txt
echo "== sweep B"; CHOWDSP_FFT_B200_LIB=$LIBB timeout 600 python tools/sweep.py --sizes 256,512,1024,2048,4096,8192,16384 --kinds r --bytes 2 2>&1 | grep -E "unordered" | tee $OUT/sweep_B.txt
echo "== sweep A again"; timeout 600 python tools/sweep.py --sizes 1024,4096,8192 --kinds r --bytes 2 2>&1 | grep -E "unordered" | tee $OUT/sweep_A2.txt
echo "== pytest A (all)" ; timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee $OUT/pytest_gpu.txt
