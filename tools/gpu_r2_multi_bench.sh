#!/bin/bash
# multi-GPU bench only (gpurun --gpus N): bench.py under torchrun, default line with the distributed sub-records
N=${1:-8}
OUT=gpurun_out/r2multib_$N
mkdir -p $OUT
echo "== bench torchrun"; ( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps 20 --warmup 3 ) 2>&1 | tail -6 | tee $OUT/bench_${N}gpu.json
echo "== bench torchrun tile_c=8 phase0?"; true
ls -la $OUT
