#!/bin/bash
# PCIe ceilings, default bench (e2e with linear copies), large sweep back on tile_fft_kernel, istft 10 warps, full GPU suite
TAG=${1:-r32}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pcie probe"; timeout 300 python tools/pcie_probe.py 2>&1 | tee $OUT/pcie_probe.txt
echo "== bench default"; timeout 600 python bench.py 2>&1 | tail -1 | tee $OUT/bench_default.json
echo "== huge"; timeout 300 python bench.py --workload huge --no-e2e --no-cpu 2>&1 | tail -1 | tee $OUT/bench_huge.json
echo "== istft"; timeout 300 python bench.py --workload istft --no-e2e --no-cpu 2>&1 | tail -1 | tee $OUT/bench_istft.json
echo "== pytest -m gpu (all)" ; timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee $OUT/pytest_gpu.txt
