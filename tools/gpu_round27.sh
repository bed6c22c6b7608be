#!/bin/bash
# ISTFT overlap-add: parity + bench; last_kernel plumbing; full GPU suite
TAG=${1:-r27}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest gpu" ; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 | tee $OUT/pytest.txt
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/smoke.txt
echo "== istft"; timeout 600 python bench.py --workload istft --steps 20 --no-e2e --no-cpu 2>&1 | tail -1 | tee $OUT/bench_istft.json
echo "== istft r16"; timeout 600 python bench.py --workload istft --steps 20 --no-e2e --no-cpu --tune radix32_mask=0 2>&1 | tail -1 | tee $OUT/bench_istft_r16.json
echo "== c2c16384"; timeout 600 python bench.py --workload c2c16384 --no-e2e --no-cpu 2>&1 | tail -1 | tee $OUT/bench_c2c16384.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:istft_kernel -s 3 -c 1 -f -o $OUT/prof_istft \
   python bench.py --workload istft --steps 2 --warmup 3 --no-e2e --no-cpu > $OUT/prof_istft.log 2>&1
ls -la $OUT
