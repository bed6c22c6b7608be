#!/bin/bash
# pipe_kernel: unordered complex output stored straight from registers (A) vs two-half staged drain (libB)
TAG=${1:-r39}
OUT=gpurun_out/$TAG
mkdir -p $OUT
LIBB=$PWD/chowdsp_fft_b200/lib/libB_prev.so
echo "== pytest pipelined" ; timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "pipelined_kernel or every_size or golden" 2>&1 | tail -4 | tee $OUT/pytest.txt
echo "== sweep A"; timeout 600 python tools/sweep.py --sizes 8192,16384 --kinds c --bytes 2 2>&1 | grep -E "C2C" | tee $OUT/sweep_A.txt
echo "== sweep B"; CHOWDSP_FFT_B200_LIB=$LIBB timeout 600 python tools/sweep.py --sizes 8192,16384 --kinds c --bytes 2 2>&1 | grep -E "C2C" | tee $OUT/sweep_B.txt
echo "== sweep A again"; timeout 600 python tools/sweep.py --sizes 8192,16384 --kinds c --bytes 2 2>&1 | grep -E "C2C" | tee $OUT/sweep_A2.txt
