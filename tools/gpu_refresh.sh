#!/bin/bash
# Regenerates the round's evidence in ONE gpurun call (about 14 minutes on one B200):
#   gpurun --timeout 1800 -- bash tools/gpu_refresh.sh <tag>      then      python tools/collect_profiles.py <tag> r02
# GPU parity suite, smoke, the default bench line (every BASELINE config as a sub-record, e2e + cpu_baseline), the reference arm, the
# size sweeps, the ncu launch list of the default command and one ncu --set full capture per dominant kernel.
TAG=${1:-refresh}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $OUT/gpu.txt 2>&1
nproc > $OUT/nproc.txt
echo "== pytest -m gpu"; ( time timeout 1500 python -m pytest tests -x -q -m gpu ) 2>&1 | tail -8 | tee $OUT/pytest_gpu.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/smoke.txt
echo "== bench default"; ( time timeout 900 python bench.py ) 2>&1 | tail -5 | tee $OUT/bench_default.json
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 5 --warmup 1 2>&1 | tail -1 | tee $OUT/bench_reference_arm.json
for wl in c2c16384 c2c16384_unordered istft; do
  echo "== bench $wl"; timeout 300 python bench.py --workload $wl --no-e2e --no-cpu 2>&1 | tail -1 | tee $OUT/bench_$wl.json
done
# every cell timed alone (1 s idle before it, best of 3 groups of 20 launches) with clock / power sampling: a back-to-back sweep runs under
# the board's power cap after ~1.5 s (profiles/r02_power_cap.txt); the sustained (power-capped) figures are taken separately below
echo "== sweep sizes"; timeout 1500 python tools/sweep.py --bytes 2 --steps 20 --pause 1.0 --repeats 3 --json $OUT/sweep_sizes.json 2>&1 | grep -E "C2C|R2C|C2R" | tee $OUT/sweep_sizes.txt
echo "== sweep sustained (back to back, no pause)"; timeout 600 python tools/sweep.py --bytes 2 --steps 20 --sizes 1024,2048,4096,8192,16384 --layouts ordered,w8 2>&1 | grep -E "C2C|R2C|C2R" | tee $OUT/sweep_sustained.txt
echo "== sweep mixed"; timeout 900 python tools/sweep.py --sizes 96,160,192,288,384,480,640,768,1920,2560,9216,12288 --bytes 2 --pause 0.5 --repeats 2 --layouts ordered,w8 2>&1 | grep -E "C2C|R2C|C2R" | tee $OUT/sweep_mixed_radix.txt
echo "== sweep stft / istft / mac over frame sizes"; timeout 600 python tools/stft_sweep.py 2>&1 | grep "N=" | tee $OUT/sweep_stft.txt
echo "== sweep partitioned convolution over block sizes"; timeout 600 python tools/pconv_sweep.py 16 2>&1 | grep pconv | tee $OUT/sweep_pconv.txt
echo "== sweep large"; timeout 600 python tools/large_sweep.py 15 16 17 18 20 22 24 26 28 2>&1 | tee $OUT/sweep_large.txt
echo "== pcie"; timeout 300 python tools/pcie_probe.py 2>&1 | tee $OUT/pcie_ceiling.txt
echo "== ncu launch list (default command, headline only)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches_c2c4096.csv \
   python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-configs > $OUT/launches_bench.log 2>&1
cap() { # name, kernel regex, workload, skip
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$2 -s $4 -c 1 -f -o $OUT/prof_$1 \
     python bench.py --workload $3 --steps 2 --warmup 3 --no-e2e --no-cpu > $OUT/prof_$1.log 2>&1
}
echo "== ncu full"
cap c2c4096 fft_kernel c2c4096 3
cap c2c4096_unordered fft_kernel c2c4096_unordered 3
cap stft wpipe_kernel stft 3
cap reverb pconv_kernel reverb 3
ls -la $OUT
