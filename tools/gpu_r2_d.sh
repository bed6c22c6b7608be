#!/bin/bash
# round-2 call D: cluster kernel with L2 prefetch, compiled-caller latency, the new bench.py default line (all configs)
OUT=gpurun_out/r2d
mkdir -p $OUT
echo "== cluster on (prefetch)"; CFB_TUNE=cluster=1 timeout 600 python tools/large_sweep.py --complex-only 15 16 17 2>&1 | tee $OUT/sweep_cluster_pf.txt
echo "== cluster on (no prefetch)"; CFB_TUNE=cluster=3 timeout 600 python tools/large_sweep.py --complex-only 15 16 17 2>&1 | tee $OUT/sweep_cluster_nopf.txt
echo "== latency C caller"
gcc -std=c11 -O2 -Iinclude tests/c_caller/latency_bench.c chowdsp_fft_b200/lib/libchowdsp_fft_b200.so -Wl,-rpath,$PWD/chowdsp_fft_b200/lib -lm -o /tmp/lat
for n in 1024 4096; do
  CHOWDSP_FFT_B200_SPIN_SYNC=1 /tmp/lat $n 2000 | tee -a $OUT/latency_c.txt
  CHOWDSP_FFT_B200_SPIN_SYNC=0 /tmp/lat $n 2000 | sed 's/^/blocking_sync /' | tee -a $OUT/latency_c.txt
done
echo "== host pointer test"; timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "host_pointer" 2>&1 | tail -5 | tee $OUT/pytest_host.txt
echo "== bench default"; ( time timeout 900 python bench.py ) 2>&1 | tail -5 | tee $OUT/bench_default.json
echo "== bench reference arm"; ( time timeout 600 python bench.py --impl reference --steps 5 --warmup 1 ) 2>&1 | tail -5 | tee $OUT/bench_reference_arm.json
echo "== full gpu suite"; ( time timeout 1800 python -m pytest tests -x -q -m gpu ) 2>&1 | tail -8 | tee $OUT/pytest_gpu.txt
ls -la $OUT
