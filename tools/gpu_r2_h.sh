#!/bin/bash
# PCIe: SM-issued zero-copy traffic vs copy engines; e2e with the whole batch transformed in place over PCIe
OUT=gpurun_out/r2h
mkdir -p $OUT
echo "== zero-copy probe"; timeout 300 tools/probes/zerocopy_probe 2>&1 | tee $OUT/zerocopy_probe.txt
echo "== e2e staged (default)"; timeout 300 python bench.py --workload c2c4096 --no-cpu --steps 10 2>&1 | tail -1 | tee $OUT/bench_e2e_staged.json
echo "== e2e zero-copy"; timeout 300 python bench.py --workload c2c4096 --no-cpu --steps 10 --tune zero_copy_kb=8388608 2>&1 | tail -1 | tee $OUT/bench_e2e_zerocopy.json
echo "== e2e zero-copy unordered"; timeout 300 python bench.py --workload c2c4096_unordered --no-cpu --steps 10 --tune zero_copy_kb=8388608 2>&1 | tail -1 | tee $OUT/bench_e2e_zerocopy_unordered.json
