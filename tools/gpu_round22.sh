#!/bin/bash
# pipe kernel with unordered variants: parity + A/B vs fft_kernel
TAG=${1:-r22}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest parity" ; timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -8 | tee $OUT/pytest.txt
echo "== sweep pipe all"; timeout 600 python tools/sweep.py --sizes 8192,16384,32768 --bytes 2 --tune pipe_mask=0xffff 2>&1 | tee $OUT/sweep_pipe.txt
echo "== sweep nopipe"; timeout 600 python tools/sweep.py --sizes 8192,16384,32768 --bytes 2 --tune pipe_mask=0 2>&1 | tee $OUT/sweep_nopipe.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pipe_kernel -s 3 -c 1 -f -o $OUT/prof_c2c16384u_pipe \
   python bench.py --workload c2c16384_unordered --steps 2 --warmup 3 --no-e2e --no-cpu > $OUT/prof_c2c16384u_pipe.log 2>&1
ls -la $OUT
