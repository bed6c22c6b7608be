#!/usr/bin/env python
"""Throughput of the multi-pass (large transform) path. GPU only."""
import math, sys, os
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch
import chowdsp_fft_b200 as cf

args = [a for a in sys.argv[1:] if not a.startswith("--")]
for a in sys.argv[1:]:
    if a.startswith("--tile-c-jfast="):
        cf.set_tuning("tile_c_jfast", int(a.split("=")[1]))
        print("tile_c_jfast override:", a.split("=")[1])
    elif a.startswith("--tile-c="):
        cf.set_tuning("tile_c", int(a.split("=")[1]))
        print("tile_c override:", a.split("=")[1])
for kv in filter(None, os.environ.get("CFB_TUNE", "").split(",")):
    k, v = kv.split("=")
    cf.set_tuning(k, int(v))
    print("tune:", k, v)
sizes = [int(a) for a in args] or [15, 16, 18, 20, 22, 24, 26, 28]
st = torch.cuda.current_stream()
for is_c in ((True,) if "--complex-only" in sys.argv else (True, False)):
    for lg in sizes:
        N = 1 << lg
        nfl = 2 * N if is_c else N
        batch = max(1, (1 << 28) // nfl)
        s = cf.fft_new_setup(N, cf.FFT_COMPLEX if is_c else cf.FFT_REAL)
        x = torch.rand(batch * nfl, device="cuda") * 2 - 1
        y = torch.empty_like(x)
        for ordered in (True, False):
            f = lambda: cf.fft_transform_batched(s, x, y, batch, nfl, nfl, cf.FFT_FORWARD, ordered, st)
            f(); f(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            for _ in range(5):
                f()
            e1.record(st); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            print(f"{'C2C' if is_c else 'R2C'} N=2^{lg} batch={batch} {'ordered' if ordered else 'unordered'} {ms:.3f} ms  "
                  f"{batch * nfl * 8 / ms / 1e6:.0f} GB/s algorithmic  {batch * (5 if is_c else 2.5) * N * math.log2(N) / ms / 1e9:.2f} TFLOP/s", flush=True)
        cf.fft_destroy_setup(s)
        del x, y
