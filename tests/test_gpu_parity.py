"""GPU parity tests: the CUDA path, called through the C ABI (libchowdsp_fft_b200.so via the ctypes
mirror), against the oracle on identical inputs.

Tolerances (north star): ordered and unordered outputs within relative L2 error 1e-6 * log2(N) of the
reference; unordered outputs compared slot for slot in the reference's own layout.  The elementwise
kernels (convolve / accumulate) are compared at 1e-6 relative L2 / bit-exact.
The restated reference tests (test/test.cpp:47-62,103-118,131-232) use the reference's own absolute
margin 2e-7 * n_floats (test/test.cpp:11).
"""
import ctypes as C
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def cf():
    import chowdsp_fft_b200 as m

    assert m.device_available(), "GPU tests need a CUDA device; there is no CPU fallback"
    return m


def dev(x):
    return torch.from_numpy(np.ascontiguousarray(x, np.float32)).cuda()


def host(t):
    return t.detach().cpu().numpy()


def all_cases(o):
    cases = []
    for is_c in (True, False):
        for lg in range(4, 16):
            N = 1 << lg
            for avx in (True, False):
                W = o.simd_width(N, is_c, avx)
                if W == 0 or (avx and W != 8):
                    continue
                if (is_c and N > 16384) or (not is_c and N > 32768):
                    continue
                cases.append((N, is_c, avx, W))
    return cases


def gpu_transform(cf, x, N, is_c, avx, backward, ordered, inplace=False):
    """x: [batch, nfl] numpy -> numpy, through fft_transform_batched on device buffers."""
    nfl = 2 * N if is_c else N
    batch = x.shape[0]
    s = cf.fft_new_setup(N, cf.FFT_COMPLEX if is_c else cf.FFT_REAL, avx)
    try:
        din = dev(x)
        dout = din if inplace else torch.full_like(din, float("nan"))
        cf.fft_transform_batched(s, din, dout, batch, nfl, nfl, cf.FFT_BACKWARD if backward else cf.FFT_FORWARD, ordered)
        torch.cuda.synchronize()
        return host(dout)
    finally:
        cf.fft_destroy_setup(s)


# --------------------------------------------------------------------------------------------------
def test_every_size_kind_layout_matches_oracle(cf, oracle_mod, ref_lib):
    o = oracle_mod
    rng = np.random.default_rng(42)
    worst = 0.0
    for N, is_c, avx, W in all_cases(o):
        nfl = 2 * N if is_c else N
        batch = 5 if N <= 4096 else 2  # odd batch: ragged last CTA when several transforms share one
        x = rng.uniform(-1, 1, (batch, nfl)).astype(np.float32)
        tol = o.parity_tol(N)
        for ordered in (True, False):
            want_f = o.np_transform(x, N, is_c, W, False, ordered)
            got_f = gpu_transform(cf, x, N, is_c, avx, False, ordered)
            e = o.rel_l2(got_f, want_f)
            assert e < tol, (N, is_c, W, ordered, "forward", e)
            got_b = gpu_transform(cf, want_f, N, is_c, avx, True, ordered)
            want_b = o.np_transform(want_f, N, is_c, W, True, ordered)
            e2 = o.rel_l2(got_b, want_b)
            assert e2 < tol, (N, is_c, W, ordered, "backward", e2)
            worst = max(worst, e, e2)
            # in place (input and output may alias, chowdsp_fft.h:136)
            assert np.array_equal(gpu_transform(cf, x, N, is_c, avx, False, ordered, inplace=True), got_f)
            if ref_lib is not None:
                ref_f, _ = ref_lib.transform(x, N, is_c, False, ordered, avx)
                assert o.rel_l2(got_f, ref_f) < tol, (N, is_c, W, ordered, "vs live reference")
                ref_b, _ = ref_lib.transform(ref_f, N, is_c, True, ordered, avx)
                assert o.rel_l2(gpu_transform(cf, ref_f, N, is_c, avx, True, ordered), ref_b) < tol
    assert worst < 5e-7  # we are in fact ~1.5e-7 from float64 truth, tighter than the reference itself


def test_golden_vectors(cf, oracle_mod, golden):
    o = oracle_mod
    n = 0
    for N in (32, 64, 128, 256, 1024, 4096):
        for is_c in (False, True):
            for W in (4, 8):
                t = f"{'c' if is_c else 'r'}{N}w{W}"
                if t + "_x" not in golden.files:
                    continue
                avx = W == 8
                tol = o.parity_tol(N)
                x = golden[t + "_x"]
                assert o.rel_l2(gpu_transform(cf, x, N, is_c, avx, False, True), golden[t + "_fwd_ordered"]) < tol
                assert o.rel_l2(gpu_transform(cf, x, N, is_c, avx, False, False), golden[t + "_fwd_unordered"]) < tol
                assert o.rel_l2(gpu_transform(cf, golden[t + "_fwd_ordered"], N, is_c, avx, True, True), golden[t + "_bwd_ordered"]) < tol
                assert o.rel_l2(gpu_transform(cf, golden[t + "_fwd_unordered"], N, is_c, avx, True, False), golden[t + "_bwd_unordered"]) < tol
                # convolve on the reference's own unordered spectra
                s = cf.fft_new_setup(N, cf.FFT_COMPLEX if is_c else cf.FFT_REAL, avx)
                fu = golden[t + "_fwd_unordered"]
                a, b, ab = dev(fu[0]), dev(fu[1]), dev(golden[t + "_conv_acc_in"])
                cf.fft_convolve_unordered(s, a, b, ab, 0.5 / N)
                assert o.rel_l2(host(ab), golden[t + "_conv_out"]) < 1e-6
                cf.fft_destroy_setup(s)
                n += 1
    assert n >= 18


def test_impulses_and_tones(cf, oracle_mod):
    o = oracle_mod
    for N, is_c, avx in ((1024, True, True), (1024, False, True), (64, True, False), (64, False, False)):
        W = o.simd_width(N, is_c, avx)
        nfl = 2 * N if is_c else N
        x = np.zeros((3, nfl), np.float32)
        x[0, 0] = 1.0
        x[1, 2 if is_c else 1] = 1.0
        k0 = 5
        n = np.arange(N)
        if is_c:
            x[2, 0::2] = np.cos(2 * np.pi * k0 * n / N)
            x[2, 1::2] = np.sin(2 * np.pi * k0 * n / N)
        else:
            x[2] = np.cos(2 * np.pi * k0 * n / N)
        for ordered in (True, False):
            got = gpu_transform(cf, x, N, is_c, avx, False, ordered)
            want = o.np_transform(x, N, is_c, W, False, ordered)
            assert np.allclose(got, want, atol=3e-6 * N ** 0.5)
        spec = gpu_transform(cf, x, N, is_c, avx, False, True)[2]
        mag = np.hypot(spec[0::2], spec[1::2])
        mag[0] = abs(spec[0]) if not is_c else mag[0]
        assert int(np.argmax(mag)) == k0


def test_batched_equals_loop_of_single_calls(cf, oracle_mod):
    """fft_transform_batched == a loop of fft_transform / fft_transform_unordered, bit for bit."""
    rng = np.random.default_rng(1)
    for N, is_c in ((4096, True), (2048, False), (64, True)):
        nfl = 2 * N if is_c else N
        s = cf.fft_new_setup(N, cf.FFT_COMPLEX if is_c else cf.FFT_REAL)
        x = dev(rng.uniform(-1, 1, (7, nfl)))
        for ordered in (True, False):
            yb = torch.empty_like(x)
            cf.fft_transform_batched(s, x, yb, 7, nfl, nfl, cf.FFT_FORWARD, ordered)
            ys = torch.empty_like(x)
            for i in range(7):
                (cf.fft_transform if ordered else cf.fft_transform_unordered)(s, x[i], ys[i], None, cf.FFT_FORWARD)
            torch.cuda.synchronize()
            assert torch.equal(yb, ys)
        cf.fft_destroy_setup(s)


@pytest.mark.parametrize("mask", [0x7FFF7FFF, 0])
@pytest.mark.parametrize("N,is_c", [(512, True), (1024, True), (8192, True), (16384, True), (1024, False), (2048, False), (16384, False), (32768, False)])
def test_radix32_geometry(cf, oracle_mod, N, is_c, mask):
    """Complex lengths 2^9, 2^10, 2^13, 2^14 exist as 16- and as 32-points-per-thread kernels (the default picks
    per size and kind): both, every kind and layout, vs the oracle."""
    o = oracle_mod
    nfl = 2 * N if is_c else N
    rng = np.random.default_rng(N)
    x = rng.uniform(-1, 1, (5, nfl)).astype(np.float32)
    cf.set_tuning("radix32_mask", mask)
    try:
        for ordered in (True, False):
            f = gpu_transform(cf, x, N, is_c, True, False, ordered)
            ref = o.np_transform(x, N, is_c, 8, False, ordered)
            assert o.rel_l2(f, ref) < o.parity_tol(N)
            b = gpu_transform(cf, ref, N, is_c, True, True, ordered, inplace=True)
            assert o.rel_l2(b, o.np_transform(ref, N, is_c, 8, True, ordered)) < o.parity_tol(N)
    finally:
        cf.set_tuning("radix32_mask", -1)


@pytest.mark.parametrize("mask", [0xFFFF, 0])
@pytest.mark.parametrize("N,is_c", [(8192, True), (16384, True), (16384, False), (32768, False)])
def test_pipelined_kernel(cf, oracle_mod, N, is_c, mask):
    """Complex lengths 2^13 / 2^14: the persistent TMA-pipelined kernel (mask 0xFFFF = every kind and layout) and
    fft_kernel (mask 0) against the oracle.  The batch exceeds the resident CTA count (so CTAs loop, with a ragged
    last round), in place and out of place; a row stride that breaks the 16-byte alignment TMA needs must fall
    back to fft_kernel and still be right."""
    o = oracle_mod
    nfl = 2 * N if is_c else N
    rng = np.random.default_rng(N + 5)
    batch = 2 * 148 * 2 + 7
    x = rng.uniform(-1, 1, (batch, nfl)).astype(np.float32)
    sub = np.r_[0:3, 147:150, 295:299, batch - 3:batch]  # first / second / third round of the persistent loop
    cf.set_tuning("pipe_mask", mask)
    try:
        for ordered in (True, False):
            ref = o.np_transform(x[sub], N, is_c, 8, False, ordered)
            f = gpu_transform(cf, x, N, is_c, True, False, ordered)
            assert o.rel_l2(f[sub], ref) < o.parity_tol(N), ordered
            assert np.array_equal(gpu_transform(cf, x, N, is_c, True, False, ordered, inplace=True), f)
            b = gpu_transform(cf, f, N, is_c, True, True, ordered)
            assert o.rel_l2(b / N, x) < o.parity_tol(N), ordered
            assert o.rel_l2(b[sub], o.np_transform(f[sub], N, is_c, 8, True, ordered)) < o.parity_tol(N), ordered
            assert np.array_equal(gpu_transform(cf, f, N, is_c, True, True, ordered, inplace=True), b)
        # rows 8 bytes off a 16-byte boundary: not a TMA source
        s = cf.fft_new_setup(N, cf.FFT_COMPLEX if is_c else cf.FFT_REAL, True)
        try:
            n0 = cf.launch_count()
            buf = torch.zeros(4 * (nfl + 2) + 2, device="cuda")
            buf[2:].as_strided((4, nfl), (nfl + 2, 1)).copy_(dev(x[:4]))
            out = torch.empty(4, nfl, device="cuda")
            cf.fft_transform_batched(s, buf[2:], out, 4, nfl + 2, nfl, cf.FFT_FORWARD, True)
            torch.cuda.synchronize()
            assert cf.launch_count() - n0 == 1
            assert o.rel_l2(host(out), o.np_transform(x[:4], N, is_c, 8, False, True)) < o.parity_tol(N)
        finally:
            cf.fft_destroy_setup(s)
    finally:
        cf.set_tuning("pipe_mask", -1)


def test_strided_batches_and_stft_gather(cf, oracle_mod):
    o = oracle_mod
    N, hop, channels, frames = 2048, 512, 3, 9
    ch_stride = (frames - 1) * hop + N + 64
    rng = np.random.default_rng(9)
    sig = rng.uniform(-1, 1, (channels, ch_stride)).astype(np.float32)
    s = cf.fft_new_setup(N, cf.FFT_REAL)
    d = dev(sig)
    out = torch.zeros(channels, frames, N, device="cuda")
    cf.fft_transform_strided(s, d, out, channels, frames, ch_stride, hop, frames * N, N, cf.FFT_FORWARD, True)
    torch.cuda.synchronize()
    got = host(out)
    for c in range(channels):
        fr = np.stack([sig[c, f * hop:f * hop + N] for f in range(frames)])
        assert o.rel_l2(got[c], o.np_transform(fr, N, False, 8, False, True)) < o.parity_tol(N)
    # padded strides in a plain batch
    x = rng.uniform(-1, 1, (4, N + 32)).astype(np.float32)
    dx = dev(x)
    dy = torch.zeros(4, N + 64, device="cuda")
    cf.fft_transform_batched(s, dx, dy, 4, N + 32, N + 64, cf.FFT_FORWARD, False)
    torch.cuda.synchronize()
    assert o.rel_l2(host(dy)[:, :N], o.np_transform(x[:, :N], N, False, 8, False, False)) < o.parity_tol(N)
    assert float(dy[:, N:].abs().max()) == 0.0  # nothing written past each transform
    cf.fft_destroy_setup(s)


@pytest.mark.parametrize("N,hop,frames", [(2048, 512, 37), (2048, 2048, 5), (512, 96, 19), (512, 130, 9), (128, 32, 18),
                                          (32, 6, 41), (32, 8, 41), (8192, 1024, 6), (32768, 4096, 3), (1024, 256, 1)])
@pytest.mark.parametrize("tail,pipe", [(6, 1), (8, 1), (8, 0)])
def test_stft_forward_window_and_layouts(cf, oracle_mod, N, hop, frames, tail, pipe):
    """fft_stft_forward == a loop of single transforms over the (windowed) overlapping frames: ordered and
    unordered, ragged last CTA group, hops that are not multiples of 4 floats (no 128-bit gather).  tail = 8 with
    pipe = 1 keeps the channels 16-byte aligned, so hops that are multiples of 4 take the persistent TMA-fed kernel
    (stft_pipe_kernel; N = 32768 does not fit its buffers and must fall back); the other two legs run stft_kernel /
    fft_kernel."""
    o = oracle_mod
    channels = 3
    samples = (frames - 1) * hop + N + tail
    cf.set_tuning("stft_pipe", pipe)
    rng = np.random.default_rng(N + hop)
    sig = rng.uniform(-1, 1, (channels, samples)).astype(np.float32)
    win = (0.5 - 0.5 * np.cos(2 * np.pi * (np.arange(N) + 0.5) / N)).astype(np.float32)
    W = o.simd_width(N, False, True)
    s = cf.fft_new_setup(N, cf.FFT_REAL)
    d, dw = dev(sig), dev(win)
    fr = np.stack([[sig[c, f * hop:f * hop + N] for f in range(frames)] for c in range(channels)]).reshape(-1, N)
    for ordered in (True, False):
        for w in (None, dw):
            out = torch.full((channels, frames, N), float("nan"), device="cuda")
            n0 = cf.launch_count()
            cf.fft_stft_forward(s, d, out, channels, frames, samples, hop, frames * N, N, w, ordered)
            torch.cuda.synchronize()
            assert cf.launch_count() - n0 == 1
            want = o.np_transform((fr * win if w is not None else fr).astype(np.float32), N, False, W, False, ordered)
            assert o.rel_l2(host(out).reshape(-1, N), want) < o.parity_tol(N), (ordered, w is not None)
    cf.set_tuning("stft_union", 1)  # the union-staging variant: same results
    try:
        for w in (None, dw):
            out = torch.full((channels, frames, N), float("nan"), device="cuda")
            cf.fft_stft_forward(s, d, out, channels, frames, samples, hop, frames * N, N, w, True)
            torch.cuda.synchronize()
            want = o.np_transform((fr * win if w is not None else fr).astype(np.float32), N, False, W, False, True)
            assert o.rel_l2(host(out).reshape(-1, N), want) < o.parity_tol(N)
    finally:
        cf.set_tuning("stft_union", 0)
        cf.set_tuning("stft_pipe", 0)
    cf.fft_destroy_setup(s)


@pytest.mark.parametrize("N,is_c,hop,frames", [(2048, False, 512, 937), (2048, False, 2048, 50), (1024, False, 100, 333), (1024, True, 2048, 700),
                                               (512, True, 1024, 5000), (2048, False, 510, 40)])
@pytest.mark.parametrize("warps", [0, 4, 1])
def test_warp_pipelined_kernel(cf, oracle_mod, N, is_c, hop, frames, warps):
    """Tuning hook wpipe: the sizes one warp owns go through wpipe_kernel (per-warp TMA prefetch pipeline) -- plain
    batches, overlapping frames, windows, ordered and unordered, more transforms than resident warps (several trips of
    every warp's loop) and fewer; hops the TMA unit cannot fetch (not 16-byte aligned) must fall back to the other
    kernels.  == a loop of single transforms."""
    o = oracle_mod
    channels = 3
    nfl = 2 * N if is_c else N
    samples = ((frames - 1) * hop + nfl + 11) // 4 * 4
    rng = np.random.default_rng(N + hop + frames)
    sig = rng.uniform(-1, 1, (channels, samples)).astype(np.float32)
    win = (0.5 - 0.5 * np.cos(2 * np.pi * (np.arange(N) + 0.5) / N)).astype(np.float32)
    W = o.simd_width(N, is_c, True)
    s = cf.fft_new_setup(N, cf.FFT_COMPLEX if is_c else cf.FFT_REAL)
    d, dw = dev(sig), dev(win)
    fr = np.stack([[sig[c, f * hop:f * hop + nfl] for f in range(frames)] for c in range(channels)]).reshape(-1, nfl)
    if warps == 0:  # the default policy: overlapping or windowed frames only
        cf.set_tuning("wpipe", -1)
        cf.fft_transform_strided(s, d, torch.empty((channels, frames, nfl), device="cuda"), channels, frames, samples, hop, frames * nfl, nfl, cf.FFT_FORWARD, True)
        assert ("wpipe_kernel" in cf.last_kernel()) == (hop % 4 == 0 and hop < nfl and nfl == 2048), cf.last_kernel()  # frames of the 2^10-point size only
    cf.set_tuning("wpipe", 1 | (warps << 8))
    try:
        for ordered in (True, False):
            for w in ((None,) if is_c else (None, dw)):
                out = torch.full((channels, frames, nfl), float("nan"), device="cuda")
                n0 = cf.launch_count()
                if w is None:
                    cf.fft_transform_strided(s, d, out, channels, frames, samples, hop, frames * nfl, nfl, cf.FFT_FORWARD, ordered)
                else:
                    cf.fft_stft_forward(s, d, out, channels, frames, samples, hop, frames * N, N, w, ordered)
                torch.cuda.synchronize()
                assert cf.launch_count() - n0 == 1
                assert ("wpipe_kernel" in cf.last_kernel()) == (hop % 4 == 0), cf.last_kernel()
                want = o.np_transform((fr * win if w is not None else fr).astype(np.float32), N, is_c, W, False, ordered)
                assert o.rel_l2(host(out).reshape(-1, nfl), want) < o.parity_tol(N), (ordered, w is not None)
    finally:
        cf.set_tuning("wpipe", -1)
    cf.fft_destroy_setup(s)


@pytest.mark.parametrize("N,hop,frames", [(2048, 512, 37), (2048, 2048, 5), (512, 96, 19), (512, 130, 9), (128, 32, 70),
                                          (32, 7, 41), (8192, 1024, 6), (16384, 4096, 3), (1024, 256, 1)])
def test_istft_overlap_add(cf, oracle_mod, N, hop, frames):
    """fft_istft_overlap_add == backward transform of every frame, window, sum at hop distance (the caller-side loop
    around the reference API, oracle.np_istft_overlap_add): ordered and unordered spectra, with and without window,
    ragged groups, odd hops; every output sample written exactly once (buffer starts as NaN, tail stays NaN)."""
    o = oracle_mod
    channels = 3
    rng = np.random.default_rng(N + hop + 9)
    x = rng.uniform(-1, 1, (channels * frames, N)).astype(np.float32)
    win = (0.5 - 0.5 * np.cos(2 * np.pi * (np.arange(N) + 0.5) / N)).astype(np.float32)
    W = o.simd_width(N, False, True)
    samples = (frames - 1) * hop + N
    s = cf.fft_new_setup(N, cf.FFT_REAL)
    dw = dev(win)
    try:
        for ordered in (True, False):
            spec = np.ascontiguousarray(o.np_transform(x, N, False, W, False, ordered).reshape(channels, frames, N))
            dspec = dev(spec)
            for w in (None, dw):
                pad = 5 if w is None else 8  # 8 keeps the channels 16-byte aligned (128-bit overlap-add path)
                out = torch.full((channels, samples + pad), float("nan"), device="cuda")
                n0 = cf.launch_count()
                cf.fft_istft_overlap_add(s, dspec, out, channels, frames, frames * N, N, samples + pad, hop, w, 1.0 / N, ordered)
                torch.cuda.synchronize()
                assert cf.launch_count() - n0 == 1
                got = host(out)
                assert np.all(np.isnan(got[:, samples:]))
                want = o.np_istft_overlap_add(spec, N, hop, W, ordered, win if w is not None else None, 1.0 / N)
                assert o.rel_l2(got[:, :samples], want) < o.parity_tol(N), (ordered, w is not None)
    finally:
        cf.fft_destroy_setup(s)


@pytest.mark.parametrize("N,hop,frames,channels", [(2048, 512, 934, 40), (2048, 1024, 300, 7), (2048, 256, 200, 3), (1024, 256, 1500, 33), (1024, 128, 77, 2), (1024, 512, 9, 1)])
@pytest.mark.parametrize("warps", [0, 3])
def test_warp_pipelined_istft(cf, oracle_mod, N, hop, frames, channels, warps):
    """Overlap-add synthesis through wistft_kernel (sizes one warp owns, hop = N/2, N/4, N/8, ordered spectra): accumulators
    in registers, segments with recomputed halos chosen by the host for the device's resident warps, several items per
    warp; == oracle.np_istft_overlap_add, every sample written exactly once, and bit-identical to a second run
    (owner-computes, no atomics).  Also the STFT -> ISTFT round trip with a Hann window at 75 % overlap."""
    o = oracle_mod
    rng = np.random.default_rng(N + hop + frames)
    x = rng.uniform(-1, 1, (channels * frames, N)).astype(np.float32)
    win = (0.5 - 0.5 * np.cos(2 * np.pi * (np.arange(N) + 0.5) / N)).astype(np.float32)
    samples = (frames - 1) * hop + N
    s = cf.fft_new_setup(N, cf.FFT_REAL)
    spec = np.ascontiguousarray(o.np_transform(x, N, False, 8, False, True).reshape(channels, frames, N))
    dspec, dw = dev(spec), dev(win)
    cf.set_tuning("wistft", 3 | (warps << 8))  # bit 1: the 2^9-point size too (it goes to ristft_kernel by default)
    try:
        for w in (None, dw):
            out = torch.full((channels, samples + 6), float("nan"), device="cuda")
            n0 = cf.launch_count()
            cf.fft_istft_overlap_add(s, dspec, out, channels, frames, frames * N, N, samples + 6, hop, w, 1.0 / N, True)
            torch.cuda.synchronize()
            assert cf.launch_count() - n0 == 1 and "wistft_kernel" in cf.last_kernel(), cf.last_kernel()
            got = host(out)
            assert np.all(np.isnan(got[:, samples:]))
            want = o.np_istft_overlap_add(spec, N, hop, 8, True, win if w is not None else None, 1.0 / N)
            assert o.rel_l2(got[:, :samples], want) < o.parity_tol(N), w is not None
            out2 = torch.full_like(out, float("nan"))
            cf.fft_istft_overlap_add(s, dspec, out2, channels, frames, frames * N, N, samples + 6, hop, w, 1.0 / N, True)
            torch.cuda.synchronize()
            assert torch.equal(out[:, :samples], out2[:, :samples])
        cf.set_tuning("wistft", 0)  # the CTA-per-segment kernel gives the same signal
        out3 = torch.full((channels, samples + 6), float("nan"), device="cuda")
        cf.fft_istft_overlap_add(s, dspec, out3, channels, frames, frames * N, N, samples + 6, hop, dw, 1.0 / N, True)
        torch.cuda.synchronize()
        assert "istft_kernel" in cf.last_kernel() and "wistft" not in cf.last_kernel()
        assert o.rel_l2(host(out3)[:, :samples], host(out)[:, :samples]) < 1e-6
    finally:
        cf.set_tuning("wistft", -1)
        cf.fft_destroy_setup(s)


@pytest.mark.parametrize("N,hop,frames,channels", [(128, 32, 900, 41), (256, 128, 300, 7), (512, 64, 210, 5), (2048, 512, 120, 9), (4096, 1024, 70, 6), (8192, 1024, 33, 3), (8192, 4096, 9, 2)])
def test_register_overlap_add_istft(cf, oracle_mod, N, hop, frames, channels):
    """Overlap-add synthesis through ristft_kernel (hop = N/2, N/4, N/8; N = 128 .. 8192; sums in the registers of the transform's own
    threads; unordered spectra at every size, ordered ones where no warp-pipelined kernel exists): == oracle.np_istft_overlap_add, every
    sample written exactly once, bit-identical between runs, and the same signal as istft_kernel (tuning hook ristft = 0)."""
    o = oracle_mod
    rng = np.random.default_rng(N + hop + frames)
    x = rng.uniform(-1, 1, (channels * frames, N)).astype(np.float32)
    win = (0.5 - 0.5 * np.cos(2 * np.pi * (np.arange(N) + 0.5) / N)).astype(np.float32)
    samples = (frames - 1) * hop + N
    W = o.simd_width(N, False, True)
    s = cf.fft_new_setup(N, cf.FFT_REAL)
    dw = dev(win)
    try:
        for ordered in (True, False):
            spec = np.ascontiguousarray(o.np_transform(x, N, False, W, False, ordered).reshape(channels, frames, N))
            dspec = dev(spec)
            for w in (None, dw):
                out = torch.full((channels, samples + 6), float("nan"), device="cuda")
                n0 = cf.launch_count()
                cf.fft_istft_overlap_add(s, dspec, out, channels, frames, frames * N, N, samples + 6, hop, w, 1.0 / N, ordered)
                torch.cuda.synchronize()
                assert cf.launch_count() - n0 == 1
                if not (ordered and N == 2048):
                    assert "ristft_kernel" in cf.last_kernel(), cf.last_kernel()
                got = host(out)
                assert np.all(np.isnan(got[:, samples:])) and not np.any(np.isnan(got[:, :samples]))
                want = o.np_istft_overlap_add(spec, N, hop, W, ordered, win if w is not None else None, 1.0 / N)
                assert o.rel_l2(got[:, :samples], want) < o.parity_tol(N), (ordered, w is not None)
                out2 = torch.full_like(out, float("nan"))
                cf.fft_istft_overlap_add(s, dspec, out2, channels, frames, frames * N, N, samples + 6, hop, w, 1.0 / N, ordered)
                torch.cuda.synchronize()
                assert torch.equal(out[:, :samples], out2[:, :samples])
            cf.set_tuning("ristft", 0)
            cf.set_tuning("wistft", 0)
            out3 = torch.full((channels, samples + 6), float("nan"), device="cuda")
            cf.fft_istft_overlap_add(s, dspec, out3, channels, frames, frames * N, N, samples + 6, hop, dw, 1.0 / N, ordered)
            torch.cuda.synchronize()
            assert "cfb::istft_kernel" in cf.last_kernel(), cf.last_kernel()
            assert o.rel_l2(host(out3)[:, :samples], host(out)[:, :samples]) < 1e-6
            cf.set_tuning("ristft", -1)
            cf.set_tuning("wistft", -1)
    finally:
        cf.set_tuning("ristft", -1)
        cf.set_tuning("wistft", -1)
        cf.fft_destroy_setup(s)


def test_istft_rejects_what_does_not_fit(cf):
    s = cf.fft_new_setup(32768, cf.FFT_REAL)
    try:
        buf = torch.zeros(3 * 32768, device="cuda")
        with pytest.raises(cf.FFTError):
            cf.fft_istft_overlap_add(s, buf, buf, 1, 2, 0, 32768, 65536, 8192, None, 1.0, True)
        with pytest.raises(cf.FFTError):
            cf.fft_istft_overlap_add(s, buf, buf, 1, 2, 0, 32768, 65536, 0, None, 1.0, True)
    finally:
        cf.fft_destroy_setup(s)


def test_stft_istft_round_trip_at_config3_size(cf, oracle_mod):
    """Size-independent property at BASELINE configs[2]'s full size per channel (480000 samples, N=2048, hop 512; 64
    channels here): analysis with a Hann window, synthesis with the same window and scale 1/N, divided by the
    window-overlap sum, returns the signal (wherever 4 frames overlap).  Few channels, so the synthesis kernel
    splits every channel into segments with recomputed halos."""
    N, hop, channels, samples = 2048, 512, 64, 480000
    frames = (samples - N) // hop + 1
    gen = torch.Generator(device="cuda").manual_seed(7)
    x = torch.rand(channels, samples, device="cuda", generator=gen) * 2 - 1
    n = torch.arange(N, device="cuda", dtype=torch.float64)
    win = (0.5 - 0.5 * torch.cos(2 * np.pi * (n + 0.5) / N)).float()
    s = cf.fft_new_setup(N, cf.FFT_REAL)
    try:
        spec = torch.empty(channels, frames, N, device="cuda")
        cf.fft_stft_forward(s, x, spec, channels, frames, samples, hop, frames * N, N, win, False)
        y = torch.zeros(channels, samples, device="cuda")
        cf.fft_istft_overlap_add(s, spec, y, channels, frames, frames * N, N, samples, hop, win, 1.0 / N, False)
        torch.cuda.synchronize()
        covered = (frames - 1) * hop + N
        wsum = torch.zeros(covered, device="cuda", dtype=torch.float64)
        w2 = (win.double() ** 2)
        for f in range(frames):
            wsum[f * hop:f * hop + N] += w2
        lo, hi = N, covered - N  # fully overlapped region
        rec = y[:, lo:hi].double() / wsum[lo:hi]
        err = torch.linalg.norm(rec - x[:, lo:hi].double()) / torch.linalg.norm(x[:, lo:hi].double())
        assert float(err) < 2 * oracle_mod.parity_tol(N)
    finally:
        cf.fft_destroy_setup(s)


def test_convolve_and_accumulate(cf, oracle_mod, ref_lib):
    o = oracle_mod
    rng = np.random.default_rng(5)
    for N, is_c, avx in ((8192, False, True), (4096, True, True), (64, False, False), (32, True, False), (256, False, False)):
        W = o.simd_width(N, is_c, avx)
        nfl = 2 * N if is_c else N
        s = cf.fft_new_setup(N, cf.FFT_COMPLEX if is_c else cf.FFT_REAL, avx)
        a = rng.uniform(-1, 1, (6, nfl)).astype(np.float32)
        b = rng.uniform(-1, 1, (6, nfl)).astype(np.float32)
        ab = rng.uniform(-1, 1, (6, nfl)).astype(np.float32)  # pre-loaded accumulator: it must ACCUMULATE
        want = o.np_convolve(a, b, ab, N, is_c, W, 0.5 / N)
        da, db, dab = dev(a), dev(b), dev(ab)
        cf.fft_convolve_unordered_batched(s, da, db, dab, 6, nfl, nfl, nfl, 0.5 / N)
        torch.cuda.synchronize()
        assert o.rel_l2(host(dab), want) < 1e-6
        if ref_lib is not None:
            assert o.rel_l2(host(dab), ref_lib.convolve(a, b, ab, N, is_c, 0.5 / N, avx)) < 1e-6
        # shared operand (stride 0) and the single-call form; aliasing ab == a is allowed (chowdsp_fft.h:153)
        dab2 = dev(ab)
        cf.fft_convolve_unordered_batched(s, da, db, dab2, 6, nfl, 0, nfl, 0.25)
        torch.cuda.synchronize()
        assert o.rel_l2(host(dab2), o.np_convolve(a, b[0][None], ab, N, is_c, W, 0.25)) < 1e-6
        alias = dev(a[0])
        cf.fft_convolve_unordered(s, alias, db[0], alias, 1.0)
        assert o.rel_l2(host(alias), o.np_convolve(a[0], b[0], a[0], N, is_c, W, 1.0)) < 1e-6
        # accumulate: plain sum, exact
        dsum = torch.empty_like(da[0])
        cf.fft_accumulate(s, da[0], db[0], dsum, nfl)
        assert np.array_equal(host(dsum), a[0] + b[0])
        cf.fft_destroy_setup(s)


def test_reference_test_suite_restated(cf, oracle_mod):
    """test/test.cpp:234-304 restated: sizes 2^5..2^15 (single-kernel range), {complex, real} x
    {malloc'd, pre-allocated} x {SSE-layout, AVX-layout}; in-place ordered fwd / bwd / scale; the
    unordered fwd -> convolve -> bwd (+ accumulate) chain.  Truth = oracle (float64), reference margin.
    Sizes above 2^14 complex / 2^15 real run through the multi-pass path."""
    o = oracle_mod
    for is_c in (True, False):
        for lg in range(5, 20):
            N = 1 << lg
            nfl = 2 * N if is_c else N
            margin = 2.0e-7 * nfl
            for avx in (False, True):
                for prealloc in (False, True):
                    W = o.simd_width(N, is_c, avx)
                    tr = cf.FFT_COMPLEX if is_c else cf.FFT_REAL
                    if prealloc:
                        nbytes = cf.fft_bytes_required(N, tr, avx)
                        block = cf.aligned_malloc(nbytes)
                        s = cf.fft_new_setup_preallocated(N, tr, block, avx)
                    else:
                        s = cf.fft_new_setup(N, tr, avx)
                    assert cf.fft_simd_width_bytes(s) == 4 * W
                    if not avx:
                        assert cf.fft_simd_width_bytes(s) == 16  # test.cpp:44-45
                    sig = o.ref_signal(N, is_c, 100.0)
                    data = cf.aligned_array(nfl)
                    work = cf.aligned_array(nfl)
                    data[:] = sig
                    cf.fft_transform(s, data, data, work, cf.FFT_FORWARD)
                    want = o.np_transform(sig, N, is_c, W, False, True)
                    assert np.max(np.abs(data - want)) <= margin
                    cf.fft_transform(s, data, data, work, cf.FFT_BACKWARD)
                    assert np.max(np.abs(data / N - sig)) <= margin
                    if not prealloc and (lg <= 12 or lg in (16, 19)):
                        sig2 = o.ref_signal(N, is_c, 200.0)
                        s1, s2, out = cf.aligned_array(nfl), cf.aligned_array(nfl), cf.aligned_array(nfl)
                        s1[:], s2[:], out[:] = sig, sig2, 0.0
                        cf.fft_transform_unordered(s, s1, s1, work, cf.FFT_FORWARD)
                        cf.fft_transform_unordered(s, s2, s2, work, cf.FFT_FORWARD)
                        cf.fft_convolve_unordered(s, s1, s2, out, 1.0 / N)
                        cf.fft_transform_unordered(s, out, out, work, cf.FFT_BACKWARD)
                        f1 = o.np_transform(sig, N, is_c, W, False, False)
                        f2 = o.np_transform(sig2, N, is_c, W, False, False)
                        want = o.np_transform(o.np_convolve(f1, f2, np.zeros(nfl, np.float32), N, is_c, W, 1.0 / N), N, is_c, W, True, False)
                        if not is_c:
                            cf.fft_accumulate(s, out, s1, out, N)  # test.cpp:218
                            want = want + f1
                        assert np.max(np.abs(out - want)) <= margin * max(1.0, float(np.max(np.abs(want))))
                        for arr in (s1, s2, out):
                            cf.aligned_free(arr.ctypes.data)
                    cf.aligned_free(data.ctypes.data)
                    cf.aligned_free(work.ctypes.data)
                    if prealloc:
                        cf.aligned_free(block)  # no fft_destroy_setup for pre-allocated plans (chowdsp_fft.h:98-113)
                    else:
                        cf.fft_destroy_setup(s)


@pytest.mark.parametrize("N", [96, 192, 384, 480, 640, 768, 9216, 1920, 24576])
def test_mixed_radix_sizes(cf, oracle_mod, ref_lib, N):
    """N = 2^a 3^b 5^c that are not powers of two (the reference's "Other sizes" tests, test/test.cpp:279-285, plus
    the largest real size of the generic kernel): every kind and layout through the batched entry point vs the
    oracle and the live reference, in place and out of place; then test_fft_complex / test_fft_real restated
    (in-place ordered forward, backward, scale; reference margin) through the reference-shaped host-pointer API."""
    o = oracle_mod
    rng = np.random.default_rng(N)
    for is_c in (True, False):
        if is_c and N > 12288:
            with pytest.raises(cf.FFTError):
                cf.fft_new_setup(N, cf.FFT_COMPLEX, True)
            continue
        nfl = 2 * N if is_c else N
        for avx in (True, False):
            W = o.simd_width(N, is_c, avx)
            assert W in (4, 8)
            x = rng.uniform(-1, 1, (5, nfl)).astype(np.float32)
            tol = o.parity_tol(N)
            for ordered in (True, False):
                want_f = o.np_transform(x, N, is_c, W, False, ordered)
                got_f = gpu_transform(cf, x, N, is_c, avx, False, ordered)
                assert o.rel_l2(got_f, want_f) < tol, (is_c, W, ordered, "forward")
                assert np.array_equal(gpu_transform(cf, x, N, is_c, avx, False, ordered, inplace=True), got_f)
                got_b = gpu_transform(cf, want_f, N, is_c, avx, True, ordered)
                assert o.rel_l2(got_b, o.np_transform(want_f, N, is_c, W, True, ordered)) < tol, (is_c, W, ordered, "backward")
                if ref_lib is not None:
                    ref_f, _ = ref_lib.transform(x, N, is_c, False, ordered, avx)
                    assert o.rel_l2(got_f, ref_f) < tol, (is_c, W, ordered, "vs live reference")
            # test/test.cpp:34-62, 92-118 restated
            s = cf.fft_new_setup(N, cf.FFT_COMPLEX if is_c else cf.FFT_REAL, avx)
            assert cf.fft_simd_width_bytes(s) == 4 * W
            sig = o.ref_signal(N, is_c, 100.0)
            data, work = cf.aligned_array(nfl), cf.aligned_array(nfl)
            data[:] = sig
            cf.fft_transform(s, data, data, work, cf.FFT_FORWARD)
            margin = 2.0e-7 * nfl
            assert np.max(np.abs(data - o.np_transform(sig, N, is_c, W, False, True))) <= margin
            cf.fft_transform(s, data, data, work, cf.FFT_BACKWARD)
            assert np.max(np.abs(data / N - sig)) <= margin
            cf.aligned_free(data.ctypes.data)
            cf.aligned_free(work.ctypes.data)
            cf.fft_destroy_setup(s)


@pytest.mark.parametrize("N,hook", [(400, None), (864, None), (2400, None), (768, 0), (9216, 0), (160, None), (288, None), (2560, None)])
def test_mixed_radix_kernel_routing(cf, oracle_mod, ref_lib, N, hook):
    """Sizes Q 2^p with Q in {3, 5, 9, 15} run in mixq_kernel, the other odd parts (25, 27, 75 ...) and everything with the
    tuning hook mixq = 0 in the generic mixed_kernel: both kernels against the oracle and the live reference, every kind and
    layout, ragged batches (a CTA holds several small transforms), and the routing itself (fft_b200_last_kernel)."""
    o = oracle_mod
    rng = np.random.default_rng(N + 5)
    try:
        if hook is not None:
            cf.set_tuning("mixq", hook)
        for is_c in (True, False):
            M = N if is_c else N // 2
            odd = M
            while odd % 2 == 0:
                odd //= 2
            W = o.simd_width(N, is_c, True)
            if W == 0:
                continue
            batch = 11
            nfl = 2 * N if is_c else N
            x = rng.uniform(-1, 1, (batch, nfl)).astype(np.float32)
            tol = o.parity_tol(N)
            for ordered in (True, False):
                want_f = o.np_transform(x, N, is_c, W, False, ordered)
                got_f = gpu_transform(cf, x, N, is_c, True, False, ordered)
                expect_mixq = hook != 0 and odd in (3, 5, 9, 15) and M // odd >= 16
                assert ("mixq_kernel" in cf.last_kernel()) == expect_mixq, (cf.last_kernel(), N, is_c)
                assert o.rel_l2(got_f, want_f) < tol, (is_c, ordered, "forward")
                got_b = gpu_transform(cf, want_f, N, is_c, True, True, ordered)
                assert o.rel_l2(got_b, o.np_transform(want_f, N, is_c, W, True, ordered)) < tol, (is_c, ordered, "backward")
                if ref_lib is not None:
                    ref_f, _ = ref_lib.transform(x, N, is_c, False, ordered, True)
                    if not is_c and N % 25 == 0:
                        # REFERENCE BUG (both the SSE and the AVX build): real transforms whose length contains 5^2 (800, 1600, 2400 ...)
                        # are wrong there -- not even self-inverse -- while its complex transforms of the same lengths are right; its own
                        # tests only reach one factor of 5 (480, 640: test/test.cpp:279-285).  This library follows the DFT.
                        assert o.rel_l2(ref_f, want_f) > 0.5
                    else:
                        assert o.rel_l2(got_f, ref_f) < tol, (is_c, ordered, "vs live reference")
    finally:
        cf.set_tuning("mixq", -1)


def test_small_transform_kernel_routing(cf, oracle_mod):
    """Dense batches of 16- / 32-point complex transforms (C2C N = 16, 32; real N = 32, 64) run in fft_small_kernel (coalesced staging
    through shared-memory rows); strided batches, misaligned bases and the tuning hook small = 0 fall back to fft_kernel.  Same results
    either way (bit-identical: the same fft_core code), ragged batches, in place, every kind and layout."""
    o = oracle_mod
    rng = np.random.default_rng(16)
    for N, is_c in [(16, True), (32, True), (32, False), (64, False)]:
        nfl = 2 * N if is_c else N
        W = o.simd_width(N, is_c, False)
        s = cf.fft_new_setup(N, cf.FFT_COMPLEX if is_c else cf.FFT_REAL, False)
        try:
            batch = 1000  # not a multiple of the CTA's 128 / 256 transforms
            x = rng.uniform(-1, 1, (batch, nfl)).astype(np.float32)
            for ordered in (True, False):
                for backward in (False, True):
                    src = x if not backward else np.ascontiguousarray(o.np_transform(x, N, is_c, W, False, ordered), np.float32)
                    want = o.np_transform(src, N, is_c, W, backward, ordered)
                    d_in, d_out = dev(src), torch.full((batch, nfl), float("nan"), device="cuda")
                    direction = cf.FFT_BACKWARD if backward else cf.FFT_FORWARD
                    cf.fft_transform_batched(s, d_in, d_out, batch, nfl, nfl, direction, ordered)
                    torch.cuda.synchronize()
                    assert "fft_small_kernel" in cf.last_kernel(), cf.last_kernel()
                    got = host(d_out)
                    assert o.rel_l2(got, want) < o.parity_tol(N), (N, is_c, ordered, backward)
                    # strided batch (gap between rows): fft_kernel, bit-identical
                    gap = nfl + 8
                    d_in2 = torch.zeros((batch, gap), device="cuda")
                    d_in2[:, :nfl] = d_in
                    d_out2 = torch.full((batch, gap), float("nan"), device="cuda")
                    cf.fft_transform_batched(s, d_in2, d_out2, batch, gap, gap, direction, ordered)
                    torch.cuda.synchronize()
                    assert "fft_small_kernel" not in cf.last_kernel()
                    assert np.array_equal(host(d_out2)[:, :nfl], got)
                    assert bool(torch.isnan(d_out2[:, nfl:]).all())
                    # in place
                    d_io = d_in.clone()
                    cf.fft_transform_batched(s, d_io, d_io, batch, nfl, nfl, direction, ordered)
                    torch.cuda.synchronize()
                    assert np.array_equal(host(d_io), got)
            cf.set_tuning("small", 0)
            d_out = torch.empty((batch, nfl), device="cuda")
            cf.fft_transform_batched(s, dev(x), d_out, batch, nfl, nfl, cf.FFT_FORWARD, True)
            torch.cuda.synchronize()
            assert "fft_small_kernel" not in cf.last_kernel()
        finally:
            cf.set_tuning("small", -1)
            cf.fft_destroy_setup(s)


@pytest.mark.parametrize("N", [32, 64, 256, 1024, 2048, 4096, 16384, 32768])
def test_juce_conventions(cf, oracle_mod, N):
    """The JUCE adapter's conventions (chowdsp_fft_juce.cpp:32-86) fused into the transform kernels: perform (inverse
    scaled by 1/N), performRealOnlyForwardTransform (Nyquist as bin N/2, optional conjugate mirror),
    performRealOnlyInverseTransform; batched, in place, rows of 2N floats with a gap between rows."""
    o = oracle_mod
    rng = np.random.default_rng(N + 17)
    batch, tol = 5, o.parity_tol(N)
    sr = cf.fft_new_setup(N, cf.FFT_REAL)
    try:
        stride = 2 * N + 4
        for ignore in (True, False):
            buf = rng.uniform(-1, 1, (batch, stride)).astype(np.float32)
            d = dev(buf)
            cf.fft_juce_real_forward_batched(sr, d, batch, stride, ignore)
            torch.cuda.synchronize()
            got, want = host(d), o.np_juce_real_forward(buf, N, ignore)
            assert o.rel_l2(got[:, :N + 2], want[:, :N + 2]) < tol
            assert np.all(got[:, 1] == 0) and np.all(got[:, N + 1] == 0)
            if ignore:
                assert np.array_equal(got[:, N + 2:], buf[:, N + 2:])  # nothing else touched
            else:
                assert o.rel_l2(got[:, :2 * N], want[:, :2 * N]) < tol
                assert np.array_equal(got[:, 2 * N:], buf[:, 2 * N:])
            # and back: inverse of the forward result returns the samples
            cf.fft_juce_real_inverse_batched(sr, d, batch, stride)
            torch.cuda.synchronize()
            back = host(d)
            assert o.rel_l2(back[:, :N], buf[:, :N]) < tol
            assert o.rel_l2(back[:, :N], o.np_juce_real_inverse(want, N)[:, :N]) < tol
    finally:
        cf.fft_destroy_setup(sr)
    if N <= 16384:
        sc = cf.fft_new_setup(N, cf.FFT_COMPLEX)
        try:
            x = rng.uniform(-1, 1, (batch, 2 * N)).astype(np.float32)
            dx, dy = dev(x), torch.empty(batch, 2 * N, device="cuda")
            cf.fft_juce_perform_batched(sc, dx, dy, batch, 2 * N, 2 * N, False)
            torch.cuda.synchronize()
            f = host(dy)
            assert o.rel_l2(f, o.np_juce_perform(x, N, False)) < tol
            cf.fft_juce_perform_batched(sc, dy, dy, batch, 2 * N, 2 * N, True)  # in place
            torch.cuda.synchronize()
            assert o.rel_l2(host(dy), o.np_juce_perform(f, N, True)) < tol
            assert o.rel_l2(host(dy), x) < tol
        finally:
            cf.fft_destroy_setup(sc)


def test_host_pointer_paths(cf, oracle_mod):
    """pageable numpy memory, pinned aligned_malloc memory (zero-copy when small, staged when large)."""
    o = oracle_mod
    rng = np.random.default_rng(3)
    N = 4096
    s = cf.fft_new_setup(N, cf.FFT_COMPLEX)
    # pageable, single call, out of place and in place
    x = rng.uniform(-1, 1, 2 * N).astype(np.float32)
    y = np.zeros_like(x)
    cf.fft_transform(s, x, y, None, cf.FFT_FORWARD)
    want = o.np_transform(x, N, True, 8, False, True)
    assert o.rel_l2(y, want) < o.parity_tol(N)
    xi = x.copy()
    cf.fft_transform(s, xi, xi, None, cf.FFT_FORWARD)
    assert np.array_equal(xi, y)
    # pinned batch large enough to take the chunked staging pipeline (3 chunks of 32 MiB)
    batch = 3 * 1024 + 17
    hin, hout = cf.aligned_array(batch * 2 * N), cf.aligned_array(batch * 2 * N)
    hin[:] = rng.uniform(-1, 1, hin.size).astype(np.float32)
    cf.fft_transform_batched(s, hin, hout, batch, 2 * N, 2 * N, cf.FFT_FORWARD, True)
    d = dev(hin.reshape(batch, 2 * N))
    dout = torch.empty_like(d)
    cf.fft_transform_batched(s, d, dout, batch, 2 * N, 2 * N, cf.FFT_FORWARD, True)
    torch.cuda.synchronize()
    assert np.array_equal(hout.reshape(batch, 2 * N), host(dout))
    sel = [0, 1, batch // 2, batch - 1]
    assert o.rel_l2(hout.reshape(batch, 2 * N)[sel], o.np_transform(hin.reshape(batch, 2 * N)[sel], N, True, 8, False, True)) < o.parity_tol(N)
    cf.aligned_free(hin.ctypes.data)
    cf.aligned_free(hout.ctypes.data)
    # mixing host and device data pointers is rejected, loudly
    with pytest.raises(cf.FFTError):
        cf.fft_transform_batched(s, x, dout, 1, 2 * N, 2 * N, cf.FFT_FORWARD, True)
    cf.fft_destroy_setup(s)


def test_setup_errors(cf):
    os.environ["CHOWDSP_FFT_B200_QUIET"] = "1"
    for N, tr in ((8, cf.FFT_COMPLEX), (16, cf.FFT_REAL), (112, cf.FFT_COMPLEX), (100, cf.FFT_REAL), (0, cf.FFT_REAL), (3 << 13, cf.FFT_COMPLEX),
                  (-4, cf.FFT_COMPLEX), (1 << 29, cf.FFT_COMPLEX), (1 << 29, cf.FFT_REAL)):
        with pytest.raises(cf.FFTError):
            cf.fft_new_setup(N, tr)
    with pytest.raises(cf.FFTError):
        cf.fft_transform_batched(12345678, None, None, 1, 1, 1, 0, True)
    # transforms that do not start on an 8-byte (unordered: 16-byte) boundary are rejected on the host, not faulted on the device
    s = cf.fft_new_setup(64, cf.FFT_REAL)
    try:
        buf = torch.zeros(2048, device="cuda")
        with pytest.raises(cf.FFTError):
            cf.fft_transform_batched(s, buf[1:], buf[1024:], 1, 64, 64, cf.FFT_FORWARD, True)   # input 4 bytes off
        with pytest.raises(cf.FFTError):
            cf.fft_transform_batched(s, buf, buf[1024:], 2, 65, 64, cf.FFT_FORWARD, True)       # odd input stride
        with pytest.raises(cf.FFTError):
            cf.fft_transform_batched(s, buf, buf[1026:], 2, 64, 66, cf.FFT_FORWARD, False)      # unordered output rows 8 bytes off
        cf.fft_transform_batched(s, buf[2:], buf[1024:], 2, 66, 68, cf.FFT_FORWARD, False)      # 8 / 16-byte aligned rows are fine
        torch.cuda.synchronize()
    finally:
        cf.fft_destroy_setup(s)


# --------------------------------------------------------------------------------------------------
# BASELINE.json full sizes: size-independent properties + sampled oracle comparison
# --------------------------------------------------------------------------------------------------
def test_config2_full_size_properties(cf, oracle_mod):
    """batched C2C N=4096 x 65536 (BASELINE config 2), ordered and unordered."""
    o = oracle_mod
    N, batch = 4096, 65536
    nfl = 2 * N
    s = cf.fft_new_setup(N, cf.FFT_COMPLEX)
    g = torch.Generator(device="cuda").manual_seed(42)
    x = torch.rand(batch, nfl, device="cuda", generator=g) * 2 - 1
    y = torch.empty_like(x)
    for ordered in (True, False):
        cf.fft_transform_batched(s, x, y, batch, nfl, nfl, cf.FFT_FORWARD, ordered)
        # Parseval per transform: sum |X|^2 = N sum |x|^2 (layout independent)
        ex = (x.double() ** 2).sum(dim=1)
        ey = (y.double() ** 2).sum(dim=1)
        assert float(((ey - N * ex).abs() / (N * ex)).max()) < 1e-5
        # sampled transforms against the oracle, slot for slot
        sel = torch.tensor([0, 1, 255, 256, 32767, 65534, 65535] + list(range(1000, 1249)), device="cuda")
        want = o.np_transform(host(x[sel]), N, True, 8, False, ordered)
        assert o.rel_l2(host(y[sel]), want) < o.parity_tol(N)
        # round trip BACKWARD(FORWARD(x)) = N x
        z = torch.empty_like(x)
        cf.fft_transform_batched(s, y, z, batch, nfl, nfl, cf.FFT_BACKWARD, ordered)
        err = (z / N - x).double().norm() / x.double().norm()
        assert float(err) < o.parity_tol(N)
        del z
    # linearity on a slice: F(a x1 + b x2) = a F(x1) + b F(x2)
    x1, x2 = x[:4096], x[4096:8192]
    comb = (0.75 * x1 - 1.25 * x2).contiguous()
    yc = torch.empty_like(comb)
    cf.fft_transform_batched(s, comb, yc, 4096, nfl, nfl, cf.FFT_FORWARD, True)
    cf.fft_transform_batched(s, x, y, 8192, nfl, nfl, cf.FFT_FORWARD, True)
    lin = 0.75 * y[:4096] - 1.25 * y[4096:8192]
    assert float((yc - lin).double().norm() / lin.double().norm()) < 1e-6
    cf.fft_destroy_setup(s)


def test_config3_stft_shape(cf, oracle_mod):
    """R2C N=2048 hop 512 (BASELINE config 3) on 32 channels x 48000 samples; every frame vs oracle for
    two channels, Parseval-style energy for all."""
    o = oracle_mod
    N, hop, channels, samples = 2048, 512, 32, 48000
    frames = (samples - N) // hop + 1
    s = cf.fft_new_setup(N, cf.FFT_REAL)
    g = torch.Generator(device="cuda").manual_seed(42)
    sig = torch.rand(channels, samples, device="cuda", generator=g) * 2 - 1
    out = torch.empty(channels, frames, N, device="cuda")
    cf.fft_transform_strided(s, sig, out, channels, frames, samples, hop, frames * N, N, cf.FFT_FORWARD, True)
    torch.cuda.synchronize()
    hs = host(sig)
    for c in (0, channels - 1):
        fr = np.stack([hs[c, f * hop:f * hop + N] for f in range(frames)])
        assert o.rel_l2(host(out[c]), o.np_transform(fr, N, False, 8, False, True)) < o.parity_tol(N)
    # inverse of every frame returns the frame (x N)
    back = torch.empty_like(out)
    cf.fft_transform_batched(s, out, back, channels * frames, N, N, cf.FFT_BACKWARD, True)
    fr_all = sig.unfold(1, N, hop)
    assert float((back / N - fr_all).double().norm() / fr_all.double().norm()) < o.parity_tol(N)
    cf.fft_destroy_setup(s)


def _ir_spectra(o, ir, N, P, W):
    channels = ir.shape[0]
    B = N // 2
    h = np.zeros((channels, P, N), np.float32)
    for p in range(P):
        seg = np.zeros((channels, N), np.float32)
        seg[:, :B] = ir[:, p * B:(p + 1) * B]
        h[:, p] = o.np_transform(seg, N, False, W, False, False)
    return h


@pytest.mark.parametrize("N,avx,P", [(8192, True, 16), (256, False, 3), (1024, True, 4)])
def test_fused_partitioned_convolution(cf, oracle_mod, ref_lib, N, avx, P):
    """fft_partitioned_convolve_step (one fused kernel per block) == the reference call sequence
    forward -> P x fft_convolve_unordered -> backward per channel and block (oracle composition and, when
    present, the live reference), == the same sequence through our own unfused entry points, and == the
    direct linear convolution."""
    o = oracle_mod
    W = o.simd_width(N, False, avx)
    channels, blocks, B = 5, P + 4, N // 2
    rng = np.random.default_rng(N)
    x = rng.uniform(-1, 1, (channels, blocks * B)).astype(np.float32)
    ir = (rng.uniform(-1, 1, (channels, P * B)) / np.sqrt(P * B)).astype(np.float32)
    h = _ir_spectra(o, ir, N, P, W)
    s = cf.fft_new_setup(N, cf.FFT_REAL, avx)
    xpad = dev(np.concatenate([np.zeros((channels, B), np.float32), x], axis=1))
    dh, fdl = dev(h), torch.zeros(channels, P, N, device="cuda")
    y = torch.zeros(channels, blocks * B, device="cuda")
    # unfused sequence through the batched entry points, kept in lock step
    fdl_u, y_u = torch.zeros_like(fdl), torch.zeros_like(y)
    spec, acc, back = (torch.zeros(channels, N, device="cuda") for _ in range(3))
    for t in range(blocks):
        cf.fft_partitioned_convolve_step(s, xpad[:, t * B:].data_ptr(), xpad.shape[1], dh, P * N, fdl, P * N,
                                         y[:, t * B:].data_ptr(), y.shape[1], channels, P, t, 1.0 / N)
        cf.fft_transform_strided(s, xpad[:, t * B:].data_ptr(), spec, channels, 1, xpad.shape[1], 0, N, 0, cf.FFT_FORWARD, False)
        fdl_u[:, t % P] = spec
        acc.zero_()
        for p in range(min(t + 1, P)):
            cf.fft_convolve_unordered_batched(s, fdl_u[:, (t - p) % P].contiguous(), dh[:, p].contiguous(), acc, channels, N, N, N, 1.0 / N)
        cf.fft_transform_batched(s, acc, back, channels, N, N, cf.FFT_BACKWARD, False)
        y_u[:, t * B:(t + 1) * B] = back[:, B:]
    torch.cuda.synchronize()
    assert torch.equal(fdl, fdl_u)  # same forward kernel code path -> identical spectra
    assert o.rel_l2(host(y), host(y_u)) < 1e-6
    want_y, want_fdl = o.np_partitioned_convolve(x, h, N, P, W)
    assert o.rel_l2(host(fdl), want_fdl) < o.parity_tol(N)
    assert o.rel_l2(host(y), want_y) < 3e-6
    for c in range(channels):
        direct = np.convolve(x[c].astype(np.float64), ir[c].astype(np.float64))[:blocks * B]
        assert o.rel_l2(host(y[c]), direct) < 1e-5
    if ref_lib is not None and avx:
        ref_y, ref_fdl, _ = ref_lib.partitioned_convolve(x, h, N, P)
        assert o.rel_l2(host(y), ref_y) < 3e-6
        assert o.rel_l2(host(fdl), ref_fdl) < o.parity_tol(N)
    cf.fft_destroy_setup(s)


def test_config4_full_size_properties(cf, oracle_mod):
    """Partitioned reverb at BASELINE config 4 scale (N=8192, 16 partitions, 4096 channels): linearity in
    the input, an impulse input reproduces the IR, and sampled channels match the oracle."""
    o = oracle_mod
    N, P, channels, B = 8192, 16, 4096, 4096
    blocks = P + 2
    s = cf.fft_new_setup(N, cf.FFT_REAL)
    g = torch.Generator(device="cuda").manual_seed(7)
    ir = (torch.rand(channels, P * B, device="cuda", generator=g) * 2 - 1) * 1e-3
    seg = torch.zeros(channels * P, N, device="cuda")
    seg[:, :B] = ir.reshape(channels * P, B)
    h = torch.empty_like(seg)
    cf.fft_transform_batched(s, seg, h, channels * P, N, N, cf.FFT_FORWARD, False)
    del seg

    def run(sig):
        xpad = torch.cat([torch.zeros(channels, B, device="cuda"), sig], dim=1).contiguous()
        fdl = torch.zeros(channels, P, N, device="cuda")
        y = torch.zeros(channels, blocks * B, device="cuda")
        for t in range(blocks):
            cf.fft_partitioned_convolve_step(s, xpad[:, t * B:].data_ptr(), xpad.shape[1], h, P * N, fdl, P * N,
                                             y[:, t * B:].data_ptr(), y.shape[1], channels, P, t, 1.0 / N)
        torch.cuda.synchronize()
        return y

    g = torch.Generator(device="cuda").manual_seed(42)
    x1 = torch.rand(channels, blocks * B, device="cuda", generator=g) * 2 - 1
    x2 = torch.rand(channels, blocks * B, device="cuda", generator=g) * 2 - 1
    y1, y2, y12 = run(x1), run(x2), run(0.5 * x1 - 2.0 * x2)
    lin = 0.5 * y1 - 2.0 * y2
    assert float((y12 - lin).double().norm() / lin.double().norm()) < 2e-6
    imp = torch.zeros(channels, blocks * B, device="cuda")
    imp[:, 0] = 1.0
    yi = run(imp)
    assert float((yi[:, :P * B] - ir).double().norm() / ir.double().norm()) < 2e-6
    assert float(yi[:, P * B:].abs().max()) < 1e-7
    for c in (0, 2047, 4095):
        direct = np.convolve(host(x1[c]).astype(np.float64), host(ir[c]).astype(np.float64))[:blocks * B]
        assert o.rel_l2(host(y1[c]), direct) < 1e-5
    cf.fft_destroy_setup(s)


@pytest.mark.parametrize("is_c", [True, False])
def test_large_transforms_multi_pass(cf, oracle_mod, ref_lib, is_c):
    """Sizes beyond one CTA (two- and three-pass four-step, large_kernels.cuh): ordered and unordered,
    forward and backward, in place and out of place, against the oracle and the live reference."""
    o = oracle_mod
    rng = np.random.default_rng(17)
    sizes = [15, 16, 17, 19, 20, 21, 22] if is_c else [16, 17, 18, 20, 21, 22, 23]
    for lg in sizes:
        N = 1 << lg
        nfl = 2 * N if is_c else N
        tol = o.parity_tol(N)
        x = rng.uniform(-1, 1, (2, nfl)).astype(np.float32)
        for avx in ((True, False) if lg <= 17 else (True,)):
            W = o.simd_width(N, is_c, avx)
            for ordered in (True, False):
                want_f = o.np_transform(x, N, is_c, W, False, ordered)
                got_f = gpu_transform(cf, x, N, is_c, avx, False, ordered)
                assert o.rel_l2(got_f, want_f) < tol, (N, is_c, W, ordered, "forward")
                got_b = gpu_transform(cf, want_f, N, is_c, avx, True, ordered)
                assert o.rel_l2(got_b, o.np_transform(want_f, N, is_c, W, True, ordered)) < tol, (N, is_c, W, ordered, "backward")
                if lg <= 17:
                    assert np.array_equal(gpu_transform(cf, x, N, is_c, avx, False, ordered, inplace=True), got_f)
                if ref_lib is not None and lg <= 19 and avx:
                    ref_f, _ = ref_lib.transform(x[:1], N, is_c, False, ordered, avx)
                    assert o.rel_l2(got_f[:1], ref_f) < tol


def test_config5_single_gpu_point(cf, oracle_mod):
    """N = 2^26 complex on one GPU (the G=1 point of BASELINE config 5 at a size the oracle finishes in
    seconds): round trip, Parseval, and a sampled-bin comparison against float64."""
    o = oracle_mod
    lg = 26
    N = 1 << lg
    s = cf.fft_new_setup(N, cf.FFT_COMPLEX)
    g = torch.Generator(device="cuda").manual_seed(42)
    x = torch.rand(2 * N, device="cuda", generator=g) * 2 - 1
    y = torch.empty_like(x)
    cf.fft_transform_batched(s, x, y, 1, 2 * N, 2 * N, cf.FFT_FORWARD, True)
    ex, ey = float((x.double() ** 2).sum()), float((y.double() ** 2).sum())
    assert abs(ey - N * ex) / (N * ex) < 1e-6
    want = o.np_transform(host(x), N, True, 8, False, True)
    assert o.rel_l2(host(y), want) < o.parity_tol(N)
    z = torch.empty_like(x)
    cf.fft_transform_batched(s, y, z, 1, 2 * N, 2 * N, cf.FFT_BACKWARD, True)
    assert float((z / N - x).double().norm() / x.double().norm()) < o.parity_tol(N)
    cf.fft_destroy_setup(s)


def _tuned(cf, **kv):
    for k, v in kv.items():
        cf.set_tuning(k, v)


@pytest.mark.parametrize("lanes,policy", [(1, 1), (3, 1), (2, 0)])
def test_l2_chunked_schedules_equal_whole_array_passes(cf, oracle_mod, lanes, policy):
    """The L2-chunked schedules (csrc/large_plan.h: build_large_schedule) run the SAME tile kernels on
    chunks whose intermediate stays in a ring buffer: results must be bit-identical to the classic whole-array passes
    (tuning l2_chunk_mb = 0), which in turn match the oracle.  Small chunk sizes force many chunks, ragged last chunks and
    ring-slot reuse at sizes the oracle finishes quickly.  Complex ordered / unordered (folded into the first / last
    pass), real forward / backward (split / merge step inside the chunk for two-pass plans)."""
    o = oracle_mod
    rng = np.random.default_rng(5)
    try:
        cf.set_tuning("cluster", 0)  # this test is about the tile passes
        for lg, is_c, batch, mb in [(16, True, 7, 1), (15, True, 9, 1), (17, False, 6, 1), (22, True, 1, 2), (22, True, 2, 4), (22, False, 2, 2), (21, True, 3, 1)]:
            N = 1 << lg
            nfl = 2 * N if is_c else N
            x = rng.uniform(-1, 1, (batch, nfl)).astype(np.float32)
            for ordered in (True, False):
                _tuned(cf, l2_chunk_mb=0)
                classic_f = gpu_transform(cf, x, N, is_c, True, False, ordered)
                want_f = o.np_transform(x[:1], N, is_c, 8, False, ordered)
                assert o.rel_l2(classic_f[:1], want_f) < o.parity_tol(N)
                classic_b = gpu_transform(cf, classic_f, N, is_c, True, True, ordered)
                _tuned(cf, l2_chunk_mb=mb, l2_lanes=lanes, l2_policy=policy)
                got_f = gpu_transform(cf, x, N, is_c, True, False, ordered)
                assert "L2-chunked" in cf.last_kernel(), cf.last_kernel()
                assert np.array_equal(got_f, classic_f), (lg, is_c, batch, ordered, "forward")
                got_b = gpu_transform(cf, classic_f, N, is_c, True, True, ordered)
                assert np.array_equal(got_b, classic_b), (lg, is_c, batch, ordered, "backward")
                assert o.rel_l2(got_b / N, x) < o.parity_tol(N)
    finally:
        _tuned(cf, l2_chunk_mb=-1, l2_lanes=-1, l2_policy=-1, cluster=-1)


def sampled_dft(torch, x2, ks):
    """float64 DFT bins X[k], k in ks, of the complex signal x2 ([N, 2] float32 on the device); phases from exact int64
    arithmetic, chunked so that the temporaries stay small."""
    N = x2.shape[0]
    out = []
    step = 1 << 24
    for k in ks:
        re = im = 0.0
        for n0 in range(0, N, step):
            n = torch.arange(n0, min(N, n0 + step), device=x2.device, dtype=torch.int64)
            ang = ((n * int(k)) % N).double() * (-2.0 * np.pi / N)
            c, s = torch.cos(ang), torch.sin(ang)
            xr, xi = x2[n0:n0 + step, 0].double(), x2[n0:n0 + step, 1].double()
            re += float((xr * c - xi * s).sum())
            im += float((xr * s + xi * c).sum())
        out.append(complex(re, im))
    return np.array(out)


def test_config5_full_size_single_gpu(cf, oracle_mod):
    """BASELINE configs[4] at its STATED size on one GPU: C2C N = 2^28 (2 GiB in, 2 GiB out).  The oracle cannot hold this in
    seconds, so parity is pinned by size-independent properties plus sampled bins: 48 bins (spread over every k1 / k2 / k3
    digit of the three-pass plan, first and last bins included) against a float64 DFT with exact integer phases, Parseval,
    forward -> backward round trip, and unordered output == the closed-form permutation of the ordered output."""
    o = oracle_mod
    free, _ = torch.cuda.mem_get_info()
    if free < 14 * (1 << 30):
        pytest.skip("needs 14 GiB of free device memory")
    lg = 28
    N = 1 << lg
    s = cf.fft_new_setup(N, cf.FFT_COMPLEX)
    g = torch.Generator(device="cuda").manual_seed(42)
    x = torch.rand(2 * N, device="cuda", generator=g) * 2 - 1
    y = torch.empty_like(x)
    cf.fft_transform_batched(s, x, y, 1, 2 * N, 2 * N, cf.FFT_FORWARD, True)
    torch.cuda.synchronize()
    rng = np.random.default_rng(28)
    ks = sorted(set([0, 1, N - 1, N // 2, (1 << 9) - 1, 1 << 9, (1 << 18) + 5] + [int(k) for k in rng.integers(0, N, 41)]))
    want = sampled_dft(torch, x.view(-1, 2), ks)
    yk = host(y.view(-1, 2)[torch.tensor(ks, device="cuda")])
    got = yk[:, 0].astype(np.float64) + 1j * yk[:, 1]
    assert np.linalg.norm(got - want) / np.linalg.norm(want) < o.parity_tol(N)
    ex, ey = float((x.double() ** 2).sum()), float((y.double() ** 2).sum())
    assert abs(ey - N * ex) / (N * ex) < 1e-6
    z = torch.empty_like(x)
    cf.fft_transform_batched(s, y, z, 1, 2 * N, 2 * N, cf.FFT_BACKWARD, True)
    assert float((z / N - x).double().norm() / x.double().norm()) < o.parity_tol(N)
    # unordered output: slot u holds ordered slot map[u]; checked on a slice of the closed form (2^20 slots) to stay in seconds
    cf.fft_transform_batched(s, x, z, 1, 2 * N, 2 * N, cf.FFT_FORWARD, False)
    torch.cuda.synchronize()
    W = 8
    u = torch.from_numpy(rng.integers(0, 2 * N, 1 << 20)).cuda()
    vec, lane = u // W, u % W           # unordered vector index, lane
    kk, is_im = vec // 2, vec % 2
    b, r = kk // W, kk % W
    bins = r * (N // W) + b * W + lane  # SURVEY.md §8a-L, complex
    assert torch.equal(z[u], y[2 * bins + is_im])
    cf.fft_destroy_setup(s)


def test_misaligned_operands_are_rejected_not_faulted(cf):
    """every entry point validates base-pointer and stride alignment on the host (FFT_B200_EINVAL) instead of
    letting a vector access fault on the device (cudaErrorMisalignedAddress is sticky and would poison the context)."""
    N = 1024
    s = cf.fft_new_setup(N, cf.FFT_REAL)
    buf = torch.zeros(8 * N + 64, device="cuda")
    a, b, ab = buf[0:N], buf[N:2 * N], buf[2 * N:3 * N]
    cf.fft_convolve_unordered_batched(s, a, b, ab, 1, N, N, N, 1.0)  # aligned: fine
    for bad in (buf[1:N + 1], buf[2:N + 2]):
        with pytest.raises(cf.FFTError):
            cf.fft_convolve_unordered_batched(s, bad, b, ab, 1, N, N, N, 1.0)
        with pytest.raises(cf.FFTError):
            cf.fft_convolve_unordered_batched(s, a, b, bad, 1, N, N, N, 1.0)
        with pytest.raises(cf.FFTError):
            cf.fft_accumulate_batched(s, a, bad, ab, N)
    with pytest.raises(cf.FFTError):  # odd stride between spectra
        cf.fft_convolve_unordered_batched(s, a, b, ab, 2, N + 2, N, N, 1.0)
    with pytest.raises(cf.FFTError):  # window not 8-byte aligned
        cf.fft_istft_overlap_add(s, buf[0:4 * N], buf[4 * N:8 * N], 1, 4, 4 * N, N, 4 * N, N // 2, buf[1:N + 1], 1.0, True)
    with pytest.raises(cf.FFTError):  # unordered spectra need 16-byte frames
        cf.fft_istft_overlap_add(s, buf[2:4 * N + 2], buf[4 * N:8 * N], 1, 4, 4 * N, N, 4 * N, N // 2, None, 1.0, False)
    with pytest.raises(cf.FFTError):
        cf.fft_juce_real_inverse_batched(s, buf[1:2 * N + 3], 1, N + 2)
    # partitioned convolution: fdl / ir 16-byte aligned, windows / output 8-byte aligned
    P = 2
    win, ir, fdl, out = buf[0:N], torch.zeros(P * N + 8, device="cuda"), torch.zeros(P * N + 8, device="cuda"), torch.zeros(N, device="cuda")
    cf.fft_partitioned_convolve_step(s, win, N, ir[:P * N], 0, fdl[:P * N], P * N, out, N // 2, 1, P, 0, 1.0)
    with pytest.raises(cf.FFTError):
        cf.fft_partitioned_convolve_step(s, win, N, ir[2:P * N + 2], 0, fdl[:P * N], P * N, out, N // 2, 1, P, 0, 1.0)
    with pytest.raises(cf.FFTError):
        cf.fft_partitioned_convolve_step(s, win, N, ir[:P * N], 0, fdl[1:P * N + 1], P * N, out, N // 2, 1, P, 0, 1.0)
    with pytest.raises(cf.FFTError):
        cf.fft_partitioned_convolve_step(s, buf[1:N + 1], N, ir[:P * N], 0, fdl[:P * N], P * N, out, N // 2, 1, P, 0, 1.0)
    cf.fft_destroy_setup(s)
    # multi-pass and mixed-radix plans: 8-byte aligned bases
    for n in (1 << 16, 96):
        s = cf.fft_new_setup(n, cf.FFT_COMPLEX)
        big = torch.zeros(4 * n + 8, device="cuda")
        with pytest.raises(cf.FFTError):
            cf.fft_transform_batched(s, big[1:2 * n + 1], big[2 * n + 2:4 * n + 2], 1, 2 * n, 2 * n, cf.FFT_FORWARD, True)
        cf.fft_destroy_setup(s)
    torch.cuda.synchronize()  # the context is still healthy
    assert float(buf.sum()) == 0.0


@pytest.mark.parametrize("lg", [15, 16, 17])
def test_cluster_one_pass_transform(cf, oracle_mod, ref_lib, lg):
    """Complex transforms of 2^15 .. 2^17 points in ONE pass on a thread-block cluster (csrc/cluster_kernels.cuh: tensor-map
    TMA loads, radix-G butterfly across the CTAs through distributed shared memory): forward / backward, natural order and
    the 8-lane unordered output, in place, batches smaller / larger than the resident cluster count and not a multiple of
    it -- against the oracle, the live reference, and bit-exact against nothing else: the tile path is a different
    algorithm, so the two are compared within tolerance."""
    o = oracle_mod
    N = 1 << lg
    tol = o.parity_tol(N)
    rng = np.random.default_rng(lg)
    try:
        cf.set_tuning("cluster_min_batch", 1)
        for batch in (1, 9, 83):
            x = rng.uniform(-1, 1, (batch, 2 * N)).astype(np.float32)
            want_f = o.np_transform(x[:3], N, True, 8, False, True)
            cf.set_tuning("cluster", 1)
            got_f = gpu_transform(cf, x, N, True, True, False, True)
            if "cluster_fft_kernel" not in cf.last_kernel():
                pytest.skip("clusters of %d CTAs cannot be co-scheduled on this device: %s" % (1 << (lg - 13), cf.last_kernel()))
            assert o.rel_l2(got_f[:3], want_f) < tol, (lg, batch, "forward")
            cf.set_tuning("cluster", 0)
            tile_f = gpu_transform(cf, x, N, True, True, False, True)
            assert "tile_fft_kernel" in cf.last_kernel()
            assert o.rel_l2(got_f, tile_f) < tol
            cf.set_tuning("cluster", 1)
            got_u = gpu_transform(cf, x, N, True, True, False, False)
            assert "cluster_fft_kernel" in cf.last_kernel()
            perm = o.np_unordered_map(N, True, 8)
            assert np.array_equal(got_u, got_f[:, perm])          # unordered == the closed-form permutation of ordered, bit for bit
            assert np.array_equal(gpu_transform(cf, x, N, True, True, False, True, inplace=True), got_f)
            got_b = gpu_transform(cf, got_f, N, True, True, True, True)
            assert "cluster_fft_kernel" in cf.last_kernel()
            assert o.rel_l2(got_b / N, x) < tol, (lg, batch, "round trip")
            if ref_lib is not None and lg <= 16 and batch == 1:
                ref_f, _ = ref_lib.transform(x[:1], N, True, False, True, True)
                assert o.rel_l2(got_f[:1], ref_f) < tol
    finally:
        cf.set_tuning("cluster", -1)
        cf.set_tuning("cluster_min_batch", -1)


def test_host_pointer_stft_istft_and_strided(cf, oracle_mod):
    """the extended entry points accept HOST buffers like the reference API does (chowdsp_fft.h:138):
    fft_stft_forward / fft_transform_strided upload every channel's unique samples once (not once per overlapping frame),
    fft_istft_overlap_add stages spectra in and the signal out.  Results are bit-identical to the device-pointer calls."""
    o = oracle_mod
    N, hop, frames, ch = 2048, 512, 61, 37
    samples = (frames - 1) * hop + N
    rng = np.random.default_rng(3)
    sig = rng.uniform(-1, 1, (ch, samples + 6)).astype(np.float32)  # channel stride > span
    win = (0.5 - 0.5 * np.cos(2 * np.pi * (np.arange(N) + 0.5) / N)).astype(np.float32)
    s = cf.fft_new_setup(N, cf.FFT_REAL)
    dsig, dwin = dev(sig), dev(win)
    for ordered in (True, False):
        for w_host, w_dev in ((None, None), (win, dwin)):
            dspec = torch.empty(ch, frames, N, device="cuda")
            cf.fft_stft_forward(s, dsig, dspec, ch, frames, sig.shape[1], hop, frames * N, N, w_dev, ordered)
            torch.cuda.synchronize()
            hspec = np.full((ch, frames, N), np.nan, np.float32)
            cf.fft_stft_forward(s, sig, hspec, ch, frames, sig.shape[1], hop, frames * N, N, w_host, ordered)
            assert np.array_equal(hspec, host(dspec)), (ordered, w_host is not None)
            fr = np.stack([sig[5, f * hop:f * hop + N] * (win if w_host is not None else 1.0) for f in range(frames)]).astype(np.float32)
            assert o.rel_l2(hspec[5], o.np_transform(fr, N, False, 8, False, ordered)) < o.parity_tol(N)
            # synthesis from host spectra
            dout = torch.empty(ch, samples, device="cuda")
            cf.fft_istft_overlap_add(s, dspec, dout, ch, frames, frames * N, N, samples, hop, w_dev, 1.0 / N, ordered)
            torch.cuda.synchronize()
            hout = np.full((ch, samples + 2), np.nan, np.float32)
            cf.fft_istft_overlap_add(s, hspec, hout, ch, frames, frames * N, N, samples + 2, hop, w_host, 1.0 / N, ordered)
            assert np.array_equal(hout[:, :samples], host(dout))
            assert np.isnan(hout[:, samples:]).all()  # nothing beyond a channel's span is touched
    # plain two-level strided batch from host memory, no window
    hspec = np.full((ch, frames, N), np.nan, np.float32)
    cf.fft_transform_strided(s, sig, hspec, ch, frames, sig.shape[1], hop, frames * N, N, cf.FFT_FORWARD, True)
    dspec = torch.empty(ch, frames, N, device="cuda")
    cf.fft_transform_strided(s, dsig, dspec, ch, frames, sig.shape[1], hop, frames * N, N, cf.FFT_FORWARD, True)
    torch.cuda.synchronize()
    assert np.array_equal(hspec, host(dspec))
    with pytest.raises(cf.FFTError):  # mixing host and device buffers is rejected
        cf.fft_stft_forward(s, sig, dspec, ch, frames, sig.shape[1], hop, frames * N, N, None, True)
    cf.fft_destroy_setup(s)


def test_tile_passes_with_32_points_per_thread(cf, oracle_mod):
    """tuning hook tile_r = 1: the 512- / 1024-point tile passes of the multi-pass path with 32 complex points per thread (two
    Stockham stages, one shared-memory exchange fewer).  Same transforms within tolerance, every layout and direction."""
    o = oracle_mod
    rng = np.random.default_rng(32)
    try:
        cf.set_tuning("tile_r", 1)
        cf.set_tuning("cluster", 0)
        for lg, is_c in [(19, True), (20, True), (22, True), (21, False), (27, True)]:
            N = 1 << lg
            nfl = 2 * N if is_c else N
            x = rng.uniform(-1, 1, (2 if lg < 27 else 1, nfl)).astype(np.float32)
            for ordered in (True, False):
                got_f = gpu_transform(cf, x, N, is_c, True, False, ordered)
                if lg <= 22:
                    assert o.rel_l2(got_f[:1], o.np_transform(x[:1], N, is_c, 8, False, ordered)) < o.parity_tol(N), (lg, is_c, ordered)
                got_b = gpu_transform(cf, got_f, N, is_c, True, True, ordered)
                assert o.rel_l2(got_b / N, x) < o.parity_tol(N), (lg, is_c, ordered, "round trip")
    finally:
        cf.set_tuning("tile_r", -1)
        cf.set_tuning("cluster", -1)


def test_tma_tile_kernel_equals_tile_fft_kernel(cf, oracle_mod):
    """tuning hook tile_tma = 1: the multi-pass path through the persistent tensor-map TMA tile kernel (cp.async.bulk.tensor
    loads into two landing slots, tensor-map stores of the output image): bit-identical to tile_fft_kernel (same butterflies
    and twiddles, only the data movement differs) for two- and three-pass plans, batches, chunked schedules, both
    directions, real transforms; and within tolerance of the oracle."""
    o = oracle_mod
    rng = np.random.default_rng(77)
    try:
        cf.set_tuning("cluster", 0)
        for lg, is_c, batch, mb in [(16, True, 5, 16), (17, True, 3, 16), (18, True, 9, 1), (20, True, 2, 16), (19, False, 3, 16), (24, True, 1, 16), (24, True, 2, 4), (25, False, 1, 16), (27, True, 1, 16)]:
            N = 1 << lg
            nfl = 2 * N if is_c else N
            x = rng.uniform(-1, 1, (batch, nfl)).astype(np.float32)
            cf.set_tuning("l2_chunk_mb", mb)
            cf.set_tuning("tile_tma", 0)
            want_f = gpu_transform(cf, x, N, is_c, True, False, True)
            want_b = gpu_transform(cf, want_f, N, is_c, True, True, True)
            cf.set_tuning("tile_tma", 1)
            got_f = gpu_transform(cf, x, N, is_c, True, False, True)
            got_b = gpu_transform(cf, want_f, N, is_c, True, True, True)
            assert np.array_equal(got_f, want_f), (lg, is_c, batch, "forward")
            assert np.array_equal(got_b, want_b), (lg, is_c, batch, "backward")
            if lg <= 20:
                assert o.rel_l2(got_f[:1], o.np_transform(x[:1], N, is_c, 8, False, True)) < o.parity_tol(N)
            # unordered layouts: only the first / last pass differs (it keeps tile_fft_kernel), the others take the TMA kernel
            got_u = gpu_transform(cf, x, N, is_c, True, False, False)
            cf.set_tuning("tile_tma", 0)
            assert np.array_equal(got_u, gpu_transform(cf, x, N, is_c, True, False, False))
    finally:
        for k in ("tile_tma", "l2_chunk_mb", "cluster"):
            cf.set_tuning(k, -1)
