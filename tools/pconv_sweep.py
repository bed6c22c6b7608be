#!/usr/bin/env python
"""Throughput of the fused partitioned-convolution step (fft_partitioned_convolve_step) over block sizes: per-channel IRs of P partitions,
channels chosen so that one step streams about 2 GiB.  Algorithmic bytes per channel-block as SURVEY.md section 8d counts them for the
reverb config (window in, delay-line write, P delay-line reads, P IR reads, N/2 samples out), scaled with N.  GPU only."""
import os
import sys
import time

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch

import chowdsp_fft_b200 as cf

P = int(sys.argv[1]) if len(sys.argv) > 1 else 16
sizes = [int(a) for a in sys.argv[2:]] or [128, 256, 512, 1024, 2048, 4096, 8192]
st = torch.cuda.current_stream()
for N in sizes:
    per_block = 2 * N + 4 * N + P * 4 * N + P * 4 * N + 2 * N  # bytes: in (N/2 new samples counted as the reverb config does: N*2), fdl write, fdl reads, ir reads, out
    channels = max(256, int(2 * 2**30 / per_block))
    s = cf.fft_new_setup(N, cf.FFT_REAL)
    win = torch.rand(channels * N, device="cuda") * 2 - 1
    ir = (torch.rand(channels * P * N, device="cuda") * 2 - 1) * 1e-3
    fdl = torch.zeros(channels * P * N, device="cuda")
    out = torch.empty(channels * (N // 2), device="cuda")
    t = [0]

    def step():
        cf.fft_partitioned_convolve_step(s, win, N, ir, P * N, fdl, P * N, out, N // 2, channels, P, t[0], 1.0 / N, st)
        t[0] += 1
    for _ in range(P + 3):
        step()
    torch.cuda.synchronize()
    time.sleep(0.5)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e30
    for _ in range(2):
        torch.cuda.synchronize()
        e0.record(st)
        for _ in range(10):
            step()
        e1.record(st)
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 10)
    print(f"pconv N={N:6d} P={P} channels={channels:7d}  {best:8.4f} ms  {channels * per_block / best / 1e6:8.1f} GB/s algorithmic  {cf.last_kernel()}", flush=True)
    cf.fft_destroy_setup(s)
    del win, ir, fdl, out
