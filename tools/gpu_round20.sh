#!/bin/bash
# pipe kernel v3 (64-bit exchange 0 in the landing buffer, two-round exchange 1); libB = pipe with table real twiddles
TAG=${1:-r20}
OUT=gpurun_out/$TAG
mkdir -p $OUT
LIBB=$PWD/chowdsp_fft_b200/lib/ab/libB.so
echo "== pytest parity" ; timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -8 | tee $OUT/pytest.txt
echo "== sweep A"; timeout 600 python tools/sweep.py --sizes 8192,16384,32768 --bytes 2 2>&1 | tee $OUT/sweep_A.txt
echo "== sweep B (pipe: table rtw)"; CHOWDSP_FFT_B200_LIB=$LIBB timeout 600 python tools/sweep.py --sizes 16384,32768 --bytes 2 --kinds r 2>&1 | tee $OUT/sweep_B.txt
echo "== sweep nopipe"; timeout 600 python tools/sweep.py --sizes 8192,16384,32768 --bytes 2 --tune pipe_mask=0 2>&1 | tee $OUT/sweep_nopipe.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pipe_kernel -s 3 -c 1 -f -o $OUT/prof_c2c16384_pipe \
   python bench.py --workload c2c16384 --steps 2 --warmup 3 --no-e2e --no-cpu > $OUT/prof_c2c16384_pipe.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pipe_kernel -s 3 -c 1 -f -o $OUT/prof_c2c8192_pipe \
   python bench.py --workload c2c8192 --steps 2 --warmup 3 --no-e2e --no-cpu > $OUT/prof_c2c8192_pipe.log 2>&1
ls -la $OUT
