#!/bin/bash
# wpipe_kernel with CTA-local dynamic work distribution: parity, STFT A/B over warps per CTA, plain-batch sweep, ncu.
TAG=${1:-r28}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest wpipe" ; timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "warp_pipelined" 2>&1 | tail -8 | tee $OUT/pytest_wpipe.txt
echo "== stft baseline"; timeout 300 python bench.py --workload stft --no-e2e --no-cpu 2>&1 | tail -1 | tee $OUT/bench_stft_base.json
for w in 0 12 11 8 6 4; do
  v=$((1 + w * 256))
  echo "== stft wpipe warps=$w"; timeout 300 python bench.py --workload stft --no-e2e --no-cpu --tune wpipe=$v 2>&1 | tail -1 | tee $OUT/bench_stft_wpipe_$w.json
done
echo "== sweep wpipe"; timeout 600 python tools/sweep.py --sizes 512,1024,2048 --bytes 2 --tune wpipe=1 2>&1 | grep fwd | tee $OUT/sweep_wpipe.txt
echo "== sweep wpipe 9 warps"; timeout 600 python tools/sweep.py --sizes 512,1024,2048 --bytes 2 --tune wpipe=2305 2>&1 | grep fwd | tee $OUT/sweep_wpipe9.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wpipe_kernel -s 3 -c 1 -f -o $OUT/prof_stft_wpipe \
   python bench.py --workload stft --steps 2 --warmup 3 --no-e2e --no-cpu --tune wpipe=1 > $OUT/prof_stft_wpipe.log 2>&1
ls -la $OUT
