#!/bin/bash
# wistft_kernel (warp-pipelined overlap-add synthesis, accumulators in registers): parity, ISTFT bench A/B, ncu.
TAG=${1:-r30}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest istft/stft" ; timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "istft or stft or warp_pipelined" 2>&1 | tail -8 | tee $OUT/pytest_istft.txt
echo "== istft wistft"; timeout 300 python bench.py --workload istft --no-e2e --no-cpu 2>&1 | tail -1 | tee $OUT/bench_istft_wistft.json
for w in 8 6; do
echo "== istft wistft warps=$w"; timeout 300 python bench.py --workload istft --no-e2e --no-cpu --tune wistft=$((1 + w*256)) 2>&1 | tail -1 | tee $OUT/bench_istft_wistft_$w.json
done
echo "== istft old"; timeout 300 python bench.py --workload istft --no-e2e --no-cpu --tune wistft=0 2>&1 | tail -1 | tee $OUT/bench_istft_old.json
echo "== stft"; timeout 300 python bench.py --workload stft --no-e2e --no-cpu 2>&1 | tail -1 | tee $OUT/bench_stft.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wistft_kernel -s 3 -c 1 -f -o $OUT/prof_wistft \
   python bench.py --workload istft --steps 2 --warmup 3 --no-e2e --no-cpu > $OUT/prof_wistft.log 2>&1
echo "== pytest -m gpu (all)" ; timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee $OUT/pytest_gpu.txt
ls -la $OUT
