#!/bin/bash
TAG=${1:-r10}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi -L | tee $OUT/gpus.txt
echo "== pytest distributed" ; timeout 900 python -m pytest tests/test_gpu_distributed.py -x -q -m gpu 2>&1 | tail -8 | tee $OUT/pytest_dist.txt
echo "== bench huge 1 GPU"; timeout 600 python bench.py --workload huge --steps 10 --no-e2e 2>&1 | tail -1 | tee $OUT/bench_huge_1.json
NG=$(nvidia-smi -L | wc -l)
if [ "$NG" -ge 2 ]; then
echo "== bench huge 2 GPUs"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload huge --steps 10 --no-e2e --no-cpu 2>&1 | tail -1 | tee $OUT/bench_huge_2.json
echo "== bench default 2 GPUs"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --no-cpu 2>&1 | tail -1 | tee $OUT/bench_default_2.json
fi
