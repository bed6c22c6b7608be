#!/bin/bash
# One gpurun call: GPU parity tests, smoke, bench lines, ncu launch list + one full capture.
# Usage (from the build container):  gpurun --timeout 1800 -- bash tools/gpu_round.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $OUT/gpu.txt 2>&1
nproc > $OUT/nproc.txt; lscpu | grep -E "Model name|^CPU\(s\)|Thread|Core|Socket" >> $OUT/nproc.txt
echo "== pytest -m gpu" ; timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee $OUT/pytest_gpu.txt
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee $OUT/smoke.txt
echo "== bench (default)" ; timeout 900 python bench.py 2>&1 | tail -3 | tee $OUT/bench_default.json
for wl in c2c4096_unordered c2c1024 c2c16384 r2c2048 r2c8192; do
  echo "== bench $wl" ; timeout 300 python bench.py --workload $wl --no-e2e --no-cpu 2>&1 | tail -1 | tee $OUT/bench_$wl.json
done
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $OUT/launches.csv \
   python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > $OUT/launches_bench.log 2>&1
echo "== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fft_kernel -s 3 -c 2 -f -o $OUT/prof_c2c4096 \
   python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > $OUT/prof_bench.log 2>&1
ls -la $OUT
