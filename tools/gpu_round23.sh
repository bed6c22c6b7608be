#!/bin/bash
# persistent TMA-fed STFT kernel: parity + A/B on the STFT config, plus pipe-kernel unordered parity
TAG=${1:-r23}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest parity" ; timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -8 | tee $OUT/pytest.txt
for r in 0x7fff7fff 0; do
echo "== stft pipe radix32_mask=$r"; timeout 600 python bench.py --workload stft --steps 20 --no-e2e --no-cpu --tune radix32_mask=$r 2>&1 | tail -1 | tee $OUT/bench_stft_pipe_$r.json
echo "== stft nopipe radix32_mask=$r"; timeout 600 python bench.py --workload stft --steps 20 --no-e2e --no-cpu --tune stft_pipe=0 --tune radix32_mask=$r 2>&1 | tail -1 | tee $OUT/bench_stft_nopipe_$r.json
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stft_pipe_kernel -s 3 -c 1 -f -o $OUT/prof_stft_pipe \
   python bench.py --workload stft --steps 2 --warmup 3 --no-e2e --no-cpu > $OUT/prof_stft_pipe.log 2>&1
ls -la $OUT
