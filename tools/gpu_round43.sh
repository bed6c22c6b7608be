#!/bin/bash
# multi-GPU round (gpurun --gpus G): distributed parity, then driver-style launches of the default, STFT and huge workloads on G GPUs
TAG=${1:-r43}
G=${2:-8}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv,noheader | tee $OUT/gpus.txt
echo "== pytest distributed"; timeout 600 python -m pytest tests/test_gpu_distributed.py -x -q -m gpu 2>&1 | tail -4 | tee $OUT/pytest_dist.txt
run() { # workload, extra args
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 1000)) bench.py --gpus $G --steps 20 --warmup 3 "$@" 2>&1 | grep '^{' | tail -1
}
echo "== default x$G"; run --no-cpu | tee $OUT/bench_default_${G}gpu.json
echo "== stft x$G"; run --workload stft --no-e2e --no-cpu | tee $OUT/bench_stft_${G}gpu.json
echo "== istft x$G"; run --workload istft --no-e2e --no-cpu | tee $OUT/bench_istft_${G}gpu.json
echo "== huge x$G (peer)"; CFB_DIST_EXCHANGE=peer run --workload huge --no-e2e --no-cpu | tee $OUT/bench_huge_${G}gpu_peer.json
echo "== huge x$G (nccl)"; CFB_DIST_EXCHANGE=nccl run --workload huge --no-e2e --no-cpu | tee $OUT/bench_huge_${G}gpu_nccl.json
