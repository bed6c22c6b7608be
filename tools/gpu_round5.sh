#!/bin/bash
TAG=${1:-r06}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest -m gpu" ; timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee $OUT/pytest_gpu.txt
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/smoke.txt
for wl in c2c4096 stft reverb; do
  echo "== bench $wl" ; timeout 900 python bench.py --workload $wl 2>&1 | tail -1 | tee $OUT/bench_$wl.json
done
echo "== ncu reverb + stft"
for wl in reverb stft; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"pconv_kernel|fft_kernel" -s 25 -c 1 -f -o $OUT/prof_$wl \
   python bench.py --workload $wl --steps 2 --warmup 3 --no-e2e --no-cpu > $OUT/prof_$wl.log 2>&1
done
ls -la $OUT
