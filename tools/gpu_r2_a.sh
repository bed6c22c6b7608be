#!/bin/bash
# round-2 call A: parity of the new large path, L2 probe, chunk-size sweep, consolidated size sweep of the shipped build
OUT=gpurun_out/r2a
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $OUT/gpu.txt 2>&1
nproc > $OUT/nproc.txt
echo "== l2 probe"; timeout 300 tools/probes/l2_probe 2>&1 | tee $OUT/l2_probe.txt
echo "== pytest new large tests"; timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "l2_chunked or config5 or misaligned or large_transforms" 2>&1 | tail -15 | tee $OUT/pytest_large.txt
echo "== large sweep default"; timeout 600 python tools/large_sweep.py 15 16 18 20 22 24 26 28 2>&1 | tee $OUT/sweep_large_default.txt
for t in "l2_chunk_mb=0" "l2_chunk_mb=8" "l2_chunk_mb=32" "l2_chunk_mb=64" "l2_chunk_mb=16,l2_lanes=1" "l2_chunk_mb=16,l2_lanes=3" "l2_chunk_mb=32,l2_lanes=3" "l2_chunk_mb=16,l2_policy=0"; do
  echo "== large sweep $t"; CFB_TUNE=$t timeout 600 python tools/large_sweep.py --complex-only 16 20 24 26 28 2>&1 | tee $OUT/sweep_large_$(echo $t | tr ',=' '__').txt
done
echo "== c caller"; timeout 600 python -m pytest tests/test_c_caller.py -x -q -m gpu -k "reference_c_test" 2>&1 | tail -5 | tee $OUT/pytest_c_caller.txt
echo "== sanitizer"; ( time timeout 1500 python -m pytest tests/test_c_caller.py -x -q -m gpu -k "sanitizer" ) 2>&1 | tail -25 | tee $OUT/pytest_sanitizer.txt
echo "== sweep sizes"; timeout 900 python tools/sweep.py --bytes 2 --steps 5 --json $OUT/sweep_sizes.json 2>&1 | grep -E "C2C|R2C|C2R" | tee $OUT/sweep_sizes.txt
echo "== full gpu suite"; ( time timeout 1500 python -m pytest tests -x -q -m gpu --deselect tests/test_c_caller.py ) 2>&1 | tail -8 | tee $OUT/pytest_gpu.txt
ls -la $OUT
