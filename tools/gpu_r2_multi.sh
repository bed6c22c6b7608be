#!/bin/bash
# round-2 multi-GPU call (gpurun --gpus N): distributed parity, bench.py under torchrun, concurrent PCIe probe
N=${1:-2}
OUT=gpurun_out/r2multi_$N
mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo.txt 2>&1
echo "== dist tests"; ( time timeout 1200 python -m pytest tests/test_gpu_distributed.py -x -q -m gpu ) 2>&1 | tail -12 | tee $OUT/pytest_dist.txt
echo "== pcie concurrent"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 tools/pcie_probe_concurrent.py 2>&1 | grep -v Warning | tail -6 | tee $OUT/pcie_concurrent.txt
echo "== bench torchrun"; ( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps 10 --warmup 3 ) 2>&1 | tail -6 | tee $OUT/bench_${N}gpu.json
ls -la $OUT
