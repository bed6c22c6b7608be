#!/bin/bash
TAG=${1:-r07}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest -m gpu" ; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee $OUT/pytest_gpu.txt
echo "== large sweep"
timeout 600 python - <<'PY' 2>&1 | tee $OUT/large.txt
import torch, math, sys
sys.path.insert(0,'.')
import chowdsp_fft_b200 as cf
st=torch.cuda.current_stream()
for is_c in (True, False):
  for lg in (15,16,18,20,22,24,26,28):
    N=1<<lg
    if not is_c and lg==28: N=1<<28
    nfl=2*N if is_c else N
    total=max(nfl, 1<<28)
    batch=total//nfl
    s=cf.fft_new_setup(N, cf.FFT_COMPLEX if is_c else cf.FFT_REAL)
    x=torch.rand(batch*nfl,device='cuda')*2-1; y=torch.empty_like(x)
    for ordered in (True, False):
        f=lambda: cf.fft_transform_batched(s,x,y,batch,nfl,nfl,cf.FFT_FORWARD,ordered,st)
        f(); f(); torch.cuda.synchronize()
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(5): f()
        e1.record(st); torch.cuda.synchronize()
        ms=e0.elapsed_time(e1)/5
        gbs=batch*nfl*8/ms/1e6
        print(f"{'C2C' if is_c else 'R2C'} N=2^{lg} batch={batch} {'ordered' if ordered else 'unordered'} {ms:.3f} ms  {gbs:.0f} GB/s algorithmic  {batch*(5 if is_c else 2.5)*N*math.log2(N)/ms/1e9:.2f} TFLOP/s", flush=True)
    cf.fft_destroy_setup(s); del x,y
PY
