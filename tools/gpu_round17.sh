#!/bin/bash
# re-validation of HEAD on a fresh box: GPU tests, smoke, both bench arms, sweep, launch list, full captures
TAG=${1:-r17}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $OUT/gpu.txt 2>&1
nproc > $OUT/nproc.txt; lscpu | grep -E "Model name|^CPU\(s\)|Thread|Core|Socket" >> $OUT/nproc.txt
echo "== pytest -m gpu" ; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee $OUT/pytest_gpu.txt
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee $OUT/smoke.txt
echo "== bench reference arm" ; timeout 900 python bench.py --impl reference 2>&1 | tail -1 | tee $OUT/bench_reference.json
echo "== bench (default)" ; timeout 900 python bench.py 2>&1 | tail -1 | tee $OUT/bench_default.json
echo "== sweep" ; timeout 900 python tools/sweep.py --json $OUT/sweep.json 2>&1 | tee $OUT/sweep.txt
for wl in stft reverb; do
  echo "== bench $wl" ; timeout 600 python bench.py --workload $wl --no-e2e --no-cpu 2>&1 | tail -1 | tee $OUT/bench_$wl.json
done
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $OUT/launches.csv \
   python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > $OUT/launches_bench.log 2>&1
echo "== ncu full"
for wl in c2c4096 c2c16384 stft; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fft_kernel\|stft_kernel -s 3 -c 1 -f -o $OUT/prof_$wl \
   python bench.py --workload $wl --steps 2 --warmup 3 --no-e2e --no-cpu > $OUT/prof_$wl.log 2>&1
done
ls -la $OUT
