#!/usr/bin/env python
"""Throughput of the STFT analysis / overlap-add synthesis entry points and of the drop-in MAC over frame sizes (hop = N/4, Hann window):
fft_stft_forward, fft_istft_overlap_add, fft_convolve_unordered_batched.  Algorithmic bytes: unique signal + spectra (STFT / ISTFT),
16 bytes per float of the spectrum (MAC).  Best of 2 groups of 10 launches after a 0.5 s pause.  GPU only."""
import os
import sys
import time

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch

import chowdsp_fft_b200 as cf

for kv in filter(None, os.environ.get("CFB_TUNE", "").split(",")):
    k, v = kv.split("=")
    cf.set_tuning(k, int(v))
    print("tune:", k, v)
st = torch.cuda.current_stream()


def timed(f):
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    time.sleep(0.5)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e30
    for _ in range(2):
        torch.cuda.synchronize()
        e0.record(st)
        for _ in range(10):
            f()
        e1.record(st)
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 10)
    return best


sizes = [int(a) for a in sys.argv[1:]] or [128, 256, 512, 1024, 2048, 4096, 8192]
for N in sizes:
    hop = N // 4
    channels = int(os.environ.get("STFT_CHANNELS", "256"))
    frames = max(8, (1 << 28) // (channels * N))  # about 1 GiB of spectra
    samples = (frames - 1) * hop + N
    s = cf.fft_new_setup(N, cf.FFT_REAL)
    sig = torch.rand(channels * samples, device="cuda") * 2 - 1
    win = None if os.environ.get("STFT_NOWIN") else torch.hann_window(N, periodic=True, device="cuda").contiguous()
    spec = torch.empty(channels * frames * N, device="cuda")
    out = torch.empty(channels * samples, device="cuda")
    nbytes = 4 * (channels * samples + channels * frames * N)
    for ordered in (True, False):
        ms = timed(lambda: cf.fft_stft_forward(s, sig, spec, channels, frames, samples, hop, frames * N, N, win, ordered, st))
        k1 = cf.last_kernel()
        ms2 = timed(lambda: cf.fft_istft_overlap_add(s, spec, out, channels, frames, frames * N, N, samples, hop, win, 1.0 / N, ordered, st))
        print(f"N={N:5d} hop={hop:5d} {'ordered  ' if ordered else 'unordered'} STFT {nbytes / ms / 1e6:7.0f} GB/s ({k1})   ISTFT {nbytes / ms2 / 1e6:7.0f} GB/s ({cf.last_kernel()})", flush=True)
    a, b, ab = spec, torch.rand_like(spec), torch.zeros_like(spec)
    ms3 = timed(lambda: cf.fft_convolve_unordered_batched(s, a, b, ab, channels * frames, N, N, N, 0.5, st))
    print(f"N={N:5d} fft_convolve_unordered_batched {16 * channels * frames * N / ms3 / 1e6:7.0f} GB/s", flush=True)
    cf.fft_destroy_setup(s)
    del sig, spec, out, a, b, ab
