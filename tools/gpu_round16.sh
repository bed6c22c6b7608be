#!/bin/bash
# multi-GPU: distributed four-step with the fused peer-store exchange vs the NCCL all-to-all
TAG=${1:-r16}
OUT=gpurun_out/$TAG
mkdir -p $OUT
NG=$(nvidia-smi -L | wc -l)
nvidia-smi -L | tee $OUT/gpus.txt
nvidia-smi topo -m > $OUT/topo.txt 2>&1
echo "== pytest distributed" ; timeout 900 python -m pytest tests/test_gpu_distributed.py -x -q -m gpu 2>&1 | tail -8 | tee $OUT/pytest_dist.txt
for G in 2 4 8; do
  if [ "$NG" -ge "$G" ]; then
    for EX in peer nccl; do
      echo "== bench huge $G GPUs exchange=$EX"
      CFB_DIST_EXCHANGE=$EX timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 2951$G bench.py --gpus $G --workload huge --steps 20 --no-e2e --no-cpu 2>&1 | tail -1 | tee $OUT/bench_huge_${G}_$EX.json
    done
  fi
done
echo "== bench huge 1 GPU"; timeout 600 python bench.py --workload huge --steps 10 --no-e2e --no-cpu 2>&1 | tail -1 | tee $OUT/bench_huge_1.json
