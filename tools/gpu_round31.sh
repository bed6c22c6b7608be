#!/bin/bash
# driver-style launches at 2 GPUs: our arm and the reference arm under torchrun
TAG=${1:-r31}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for impl in reference ours; do
  echo "== torchrun 2 ranks --impl $impl"
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 10 --warmup 3 --impl $impl > $OUT/bench_2gpu_$impl.log 2>&1
  echo "rc=$?"; tail -2 $OUT/bench_2gpu_$impl.log | cut -c1-600
done
echo "== stft 2 ranks"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --workload stft --steps 10 --no-cpu 2>&1 | tail -1 | cut -c1-400 | tee $OUT/bench_stft_2gpu.json
ls -la $OUT
