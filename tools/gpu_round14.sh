#!/bin/bash
TAG=${1:-r14}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest large/dist/stft"; timeout 900 python -m pytest tests -x -q -m gpu -k "large or config5 or dist or stft or strided" 2>&1 | tail -8 | tee $OUT/pytest.txt
echo "== large sweep tile_c=8"; timeout 600 python tools/large_sweep.py --tile-c=8 18 20 24 26 28 2>&1 | tee $OUT/large_c8.txt
echo "== large sweep tile_c=16"; timeout 600 python tools/large_sweep.py --tile-c=16 18 20 24 26 28 2>&1 | tee $OUT/large_c16.txt
echo "== bench stft"; timeout 600 python bench.py --workload stft --steps 10 --no-e2e --no-cpu 2>&1 | tail -1 | tee $OUT/bench_stft.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches_huge.csv \
   python bench.py --workload huge --steps 2 --warmup 3 --no-e2e --no-cpu > $OUT/launches_huge.log 2>&1
ls -la $OUT
