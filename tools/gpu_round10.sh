#!/bin/bash
# Re-entry validation round: full GPU parity suite, smoke, default bench, large/huge numbers, launch list.
TAG=${1:-r10}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi -L | tee $OUT/gpus.txt
nproc > $OUT/nproc.txt; lscpu | grep -E "Model name|^CPU\(s\)|Thread|Core|Socket" >> $OUT/nproc.txt
echo "== pytest -m gpu" ; timeout 1500 python -m pytest tests -x -q -m gpu --durations=8 2>&1 | tail -25 | tee $OUT/pytest_gpu.txt
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee $OUT/smoke.txt
echo "== bench (default)" ; timeout 900 python bench.py 2>&1 | tail -3 | tee $OUT/bench_default.json
echo "== bench reference arm" ; timeout 900 python bench.py --impl reference --steps 5 --warmup 3 2>&1 | tail -1 | tee $OUT/bench_reference.json
for wl in stft reverb huge; do
  echo "== bench $wl" ; timeout 600 python bench.py --workload $wl --steps 10 --no-e2e --no-cpu 2>&1 | tail -1 | tee $OUT/bench_$wl.json
done
echo "== sweep"; timeout 600 python tools/sweep.py 2>&1 | tee $OUT/sweep.txt
echo "== large sweep"; timeout 600 python tools/large_sweep.py 2>&1 | tee $OUT/large.txt
echo "== ncu launch list (huge)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches_huge.csv \
   python bench.py --workload huge --steps 2 --warmup 3 --no-e2e --no-cpu > $OUT/launches_huge.log 2>&1
ls -la $OUT
