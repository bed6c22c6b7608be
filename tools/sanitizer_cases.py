#!/usr/bin/env python
"""One small case per kernel family, meant to run UNDER compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool racecheck python tools/sanitizer_cases.py
fft_kernel (ordered / unordered, real / complex), pipe_kernel (2^13, 2^14), wpipe_kernel + wistft_kernel (STFT / ISTFT),
istft_kernel, stft_kernel, pconv_kernel, mixq_kernel, mixed_kernel, tile_fft_kernel (+ the L2-prefetch variant) (classic + L2-chunked, unordered folded in),
real_pass_kernel, convolve / accumulate, the distributed phase kernels with the peer-store epilogue (world = 1).
Results are checked loosely (finite, right norm): parity proper lives in tests/.  GPU only."""
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import numpy as np
import torch

import chowdsp_fft_b200 as cf

F, B = cf.FFT_FORWARD, cf.FFT_BACKWARD
g = torch.Generator(device="cuda").manual_seed(1)


def rnd(n):
    return torch.rand(n, device="cuda", generator=g) * 2 - 1


def roundtrip(N, is_c, batch, ordered, avx=True):
    nfl = 2 * N if is_c else N
    s = cf.fft_new_setup(N, cf.FFT_COMPLEX if is_c else cf.FFT_REAL, avx)
    x, y, z = rnd(batch * nfl), torch.empty(batch * nfl, device="cuda"), torch.empty(batch * nfl, device="cuda")
    cf.fft_transform_batched(s, x, y, batch, nfl, nfl, F, ordered)
    k1 = cf.last_kernel()
    cf.fft_transform_batched(s, y, z, batch, nfl, nfl, B, ordered)
    torch.cuda.synchronize()
    err = float((z / N - x).norm() / x.norm())
    assert err < 1e-5, (N, is_c, ordered, err)
    cf.fft_destroy_setup(s)
    print(f"ok round trip N={N} {'C2C' if is_c else 'R2C'} {'ordered' if ordered else 'unordered'} avx={avx}: {k1} | {cf.last_kernel()}", flush=True)


for N, is_c in [(16, True), (32, True), (32, False), (64, False)]:  # fft_small_kernel (dense batches of tiny transforms)
    for ordered in (True, False):
        roundtrip(N, is_c, 300, ordered, avx=False)
for N, is_c in [(64, True), (1024, True), (4096, True), (2048, False), (8192, False)]:
    for ordered in (True, False):
        roundtrip(N, is_c, 3, ordered)
roundtrip(256, True, 5, False, avx=False)
for N, is_c in [(8192, True), (16384, True), (32768, False)]:   # pipe_kernel
    for ordered in (True, False):
        roundtrip(N, is_c, 3, ordered)
for N in (96, 480):                                              # mixed radix: mixq_kernel (odd part 3 / 15)
    roundtrip(N, True, 3, True)
    roundtrip(N * 4, False, 3, False)
for N in (400, 864):                                             # ... and the generic mixed_kernel (odd part 25 / 27)
    roundtrip(N, True, 3, True)
    roundtrip(N * 2, False, 3, False)
cf.set_tuning("tile_pf", 4)                                      # tile_fft_pf_kernel (tensor-map L2 prefetch)
roundtrip(1 << 16, True, 5, True)
cf.set_tuning("tile_pf", -1)
# multi-pass: classic and L2-chunked (tiny chunks), unordered folded in; real split / merge
for mb in (0, 1):
    cf.set_tuning("l2_chunk_mb", mb)
    cf.set_tuning("l2_lanes", 3)
    for N, is_c, batch in [(1 << 15, True, 9), (1 << 17, False, 9), (1 << 21, True, 1)]:
        for ordered in (True, False):
            roundtrip(N, is_c, batch, ordered)
cf.set_tuning("l2_chunk_mb", -1)
cf.set_tuning("l2_lanes", -1)

# STFT / ISTFT: wpipe_kernel + wistft_kernel (N = 2048, hop 512), stft_kernel / istft_kernel (other hops / sizes)
# (N = 512 hop 128 and N = 4096 hop 1024 synthesise through ristft_kernel, N = 512 hop 96 through istft_kernel)
for N, hop, frames, ch in [(2048, 512, 23, 3), (1024, 256, 17, 2), (512, 96, 11, 2), (512, 128, 29, 5), (4096, 1024, 9, 2)]:
    s = cf.fft_new_setup(N, cf.FFT_REAL)
    samples = (frames - 1) * hop + N
    sig = rnd(ch * samples)
    win = torch.hann_window(N, periodic=True, device="cuda").contiguous()
    spec = torch.empty(ch * frames * N, device="cuda")
    out = torch.empty(ch * samples, device="cuda")
    cf.fft_stft_forward(s, sig, spec, ch, frames, samples, hop, frames * N, N, win, True)
    k1 = cf.last_kernel()
    cf.fft_istft_overlap_add(s, spec, out, ch, frames, frames * N, N, samples, hop, win, 1.0 / N, True)
    torch.cuda.synchronize()
    assert bool(torch.isfinite(out).all())
    cf.fft_destroy_setup(s)
    print(f"ok stft/istft N={N} hop={hop}: {k1} | {cf.last_kernel()}", flush=True)

# partitioned convolution (pconv_kernel) + elementwise kernels
N, P, ch = 1024, 3, 4
s = cf.fft_new_setup(N, cf.FFT_REAL)
ir, fdl = rnd(ch * P * N) * 1e-3, torch.zeros(ch * P * N, device="cuda")
out = torch.empty(ch * (N // 2), device="cuda")
for t in range(5):
    cf.fft_partitioned_convolve_step(s, rnd(ch * N), N, ir, P * N, fdl, P * N, out, N // 2, ch, P, t, 1.0 / N)
a, b, ab = rnd(ch * N), rnd(ch * N), torch.zeros(ch * N, device="cuda")
cf.fft_convolve_unordered_batched(s, a, b, ab, ch, N, N, N, 0.5)
cf.fft_accumulate_batched(s, a, b, ab, ch * N)
torch.cuda.synchronize()
assert bool(torch.isfinite(out).all()) and bool(torch.isfinite(ab).all())
cf.fft_destroy_setup(s)
print("ok pconv / convolve / accumulate", flush=True)
print("all sanitizer cases ran", flush=True)
