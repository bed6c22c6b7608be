#!/bin/bash
# round-2 call B: cluster one-pass kernel (parity, throughput, ncu), distributed C-ABI context on 1 GPU, sanitizer + suite
OUT=gpurun_out/r2b
mkdir -p $OUT
echo "== cluster parity"; timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "cluster_one_pass or l2_chunked" 2>&1 | tail -15 | tee $OUT/pytest_cluster.txt
echo "== dist world1"; timeout 600 python -m pytest tests/test_gpu_distributed.py -x -q -m gpu -k "world1" 2>&1 | tail -8 | tee $OUT/pytest_dist1.txt
echo "== large sweep cluster on"; timeout 600 python tools/large_sweep.py --complex-only 15 16 17 18 20 22 24 2>&1 | tee $OUT/sweep_cluster_on.txt
echo "== large sweep cluster off"; CFB_TUNE=cluster=0 timeout 600 python tools/large_sweep.py --complex-only 15 16 17 2>&1 | tee $OUT/sweep_cluster_off.txt
echo "== ncu cluster"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cluster_fft_kernel -s 2 -c 1 -f -o $OUT/prof_cluster16 \
   python tools/large_sweep.py --complex-only 16 > $OUT/prof_cluster16.log 2>&1
echo "== sanitizer"; ( time timeout 1500 python -m pytest tests/test_c_caller.py -x -q -m gpu -k "sanitizer" ) 2>&1 | tail -25 | tee $OUT/pytest_sanitizer.txt
echo "== full gpu suite"; ( time timeout 1800 python -m pytest tests -x -q -m gpu --deselect tests/test_c_caller.py::test_compute_sanitizer_clean ) 2>&1 | tail -8 | tee $OUT/pytest_gpu.txt
ls -la $OUT
