#!/bin/bash
# round-2 call C: cluster kernel v2 (st.async exchange, TMA store), host-pointer STFT/ISTFT, config-1 latency
OUT=gpurun_out/r2c
mkdir -p $OUT
echo "== cluster + host pointer parity"; timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "cluster_one_pass or host_pointer" 2>&1 | tail -15 | tee $OUT/pytest_cluster.txt
echo "== large sweep cluster on"; timeout 600 python tools/large_sweep.py --complex-only 15 16 17 2>&1 | tee $OUT/sweep_cluster_on.txt
echo "== large sweep cluster off"; CFB_TUNE=cluster=0 timeout 600 python tools/large_sweep.py --complex-only 15 16 17 2>&1 | tee $OUT/sweep_cluster_off.txt
echo "== single1024"; timeout 300 python bench.py --workload single1024 2>&1 | tail -1 | tee $OUT/bench_single1024.json
echo "== single1024 no spin"; timeout 300 python bench.py --workload single1024 --tune spin_sync=0 --no-cpu 2>&1 | tail -1 | tee $OUT/bench_single1024_nospin.json
echo "== ncu cluster"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cluster_fft_kernel -s 2 -c 1 -f -o $OUT/prof_cluster16 \
   python tools/large_sweep.py --complex-only 16 > $OUT/prof_cluster16.log 2>&1
echo "== full gpu suite"; ( time timeout 1800 python -m pytest tests -x -q -m gpu ) 2>&1 | tail -8 | tee $OUT/pytest_gpu.txt
ls -la $OUT
