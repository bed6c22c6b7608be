"""TEST INFRASTRUCTURE ONLY -- the checker for the CUDA path, never the product.

Only ``tests/``, ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs and
``__graft_entry__.smoke()`` may import this module; nothing under ``chowdsp_fft_b200/`` does.

Three layers, all CPU:

* ``RefLib``       ctypes view of ``oracle/_ref/libchowdsp_fft_ref.so`` = the UNMODIFIED reference
                   (``/root/reference/chowdsp_fft.cpp`` + ``simd/chowdsp_fft_impl_avx.cpp``) built by
                   ``oracle/Makefile`` plus our thread-per-core driver ``oracle/ref_driver.cpp``.
* ``COracle``      ctypes view of ``oracle/_ref/liboracle_fft.so`` = ``oracle/oracle_fft.c``, the
                   plain-C restatement (double-precision internals).
* ``np_*``         numpy/float64 restatement of the same semantics (fast at 2^20+ points), following
                   the same reference lines as the C file:
                     np_unordered_map  <- pffft_zreorder, simd/chowdsp_fft_impl_avx.cpp:1780-1839
                     np_transform      <- pffft_transform_internal, avx:1848-1935
                     np_convolve       <- pffft_convolve_internal, avx:1937-1979
                     np_accumulate     <- fft_accumulate_internal, avx:1981-1994

Parity status: PINNED against the reference itself (tests/test_oracle.py: live ``RefLib`` when the
.so is present, and the committed ``tests/golden/*.npz`` produced from it by tests/gen_golden.py).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from functools import lru_cache

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
REF_SO = os.path.join(REF_DIR, "libchowdsp_fft_ref.so")
ORACLE_SO = os.path.join(REF_DIR, "liboracle_fft.so")

FFT_FORWARD, FFT_BACKWARD = 0, 1
FFT_REAL, FFT_COMPLEX = 0, 1

_fp = C.POINTER(C.c_float)


def build(verbose: bool = False) -> None:
    """Compile the C restatement and (when /root/reference exists) the reference .so."""
    r = subprocess.run(["make", "-C", HERE, "all"], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("oracle build failed")


def _ptr(a: np.ndarray):
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_fp)


def aligned_empty(nfloats: int, align: int = 64) -> np.ndarray:
    """fp32 array whose data pointer is `align`-byte aligned (the reference does aligned vector loads)."""
    raw = np.empty(nfloats * 4 + align, dtype=np.uint8)
    off = (-raw.ctypes.data) % align
    return raw[off:off + nfloats * 4].view(np.float32)


def aligned_copy(x, align: int = 64) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32).ravel()
    out = aligned_empty(x.size, align)
    out[:] = x
    return out


# --------------------------------------------------------------------------------------------------
# numpy restatement
# --------------------------------------------------------------------------------------------------
def simd_width(N: int, is_complex: bool, use_avx: bool = True) -> int:
    """W (floats) the reference picks: 8 = AVX handle, 4 = SSE handle, 0 = unsupported.
    chowdsp_fft.cpp:258-280 + common.hpp:168-177; N = 2^a 3^b 5^c (common.hpp:51-75 decompose)."""
    if N <= 0:
        return 0
    m = N
    for r in (2, 3, 5):
        while m % r == 0:
            m //= r
    if m != 1:
        return 0
    for W in ((8, 4) if use_avx else (4,)):
        if N % (W * W if is_complex else 2 * W * W) == 0:
            return W
    return 0


@lru_cache(maxsize=64)
def np_unordered_map(N: int, is_complex: bool, W: int) -> np.ndarray:
    """map[u] = ordered float slot stored at unordered float slot u."""
    if is_complex:
        k = np.arange(N // W)
        b, r = np.divmod(k, W)
        j = np.arange(W)
        bins = (r * (N // W) + b * W)[:, None] + j[None, :]
    else:
        Q = N // (2 * W)
        k = np.arange(Q)
        b, r = np.divmod(k, W)
        m = b[:, None] * W + np.arange(W)[None, :]
        bins = r[:, None] * Q + np.where((r % 2 == 1)[:, None], (Q - m) % Q, m)
    out = np.empty((bins.shape[0], 2, W), dtype=np.int64)
    out[:, 0, :] = 2 * bins
    out[:, 1, :] = 2 * bins + 1
    out = out.reshape(-1)
    out.setflags(write=False)
    return out


def np_transform(x, N: int, is_complex: bool, W: int, backward: bool, ordered: bool) -> np.ndarray:
    """fft_transform / fft_transform_unordered on a batch: x has shape [..., nfloats]; float64 inside."""
    x = np.asarray(x, dtype=np.float32)
    nfl = 2 * N if is_complex else N
    assert x.shape[-1] == nfl
    lead = x.shape[:-1]
    x = x.reshape(-1, nfl).astype(np.float64)
    pmap = None if ordered else np_unordered_map(N, is_complex, W)
    if not backward:
        if is_complex:
            X = np.fft.fft(x[:, 0::2] + 1j * x[:, 1::2], axis=-1)
            freq = np.empty_like(x)
            freq[:, 0::2], freq[:, 1::2] = X.real, X.imag
        else:
            X = np.fft.rfft(x, axis=-1)
            freq = np.empty_like(x)
            freq[:, 0] = X[:, 0].real
            freq[:, 1] = X[:, N // 2].real
            freq[:, 2::2], freq[:, 3::2] = X[:, 1:N // 2].real, X[:, 1:N // 2].imag
        out = freq if ordered else freq[:, pmap]
    else:
        if ordered:
            freq = x
        else:
            freq = np.empty_like(x)
            freq[:, pmap] = x
        if is_complex:
            out_c = np.fft.ifft(freq[:, 0::2] + 1j * freq[:, 1::2], axis=-1) * N
            out = np.empty_like(x)
            out[:, 0::2], out[:, 1::2] = out_c.real, out_c.imag
        else:
            X = np.zeros((x.shape[0], N // 2 + 1), dtype=np.complex128)
            X[:, 0] = freq[:, 0]
            X[:, N // 2] = freq[:, 1]
            X[:, 1:N // 2] = freq[:, 2::2] + 1j * freq[:, 3::2]
            out = np.fft.irfft(X, n=N, axis=-1) * N
    return out.astype(np.float32).reshape(*lead, nfl)


def np_convolve(a, b, ab, N: int, is_complex: bool, W: int, scaling: float) -> np.ndarray:
    """Returns ab + a*b*scaling in the unordered domain (batch on leading dims)."""
    a = np.asarray(a, np.float32).astype(np.float64)
    b = np.asarray(b, np.float32).astype(np.float64)
    ab = np.asarray(ab, np.float32).astype(np.float64)
    nfl = 2 * N if is_complex else N
    shp = np.broadcast_shapes(a.shape, b.shape, ab.shape)
    a, b, ab = (np.broadcast_to(t, shp).reshape(-1, nfl // (2 * W), 2, W) for t in (a, b, ab))
    out = np.empty(ab.shape, dtype=np.float64)
    out[:, :, 0, :] = ab[:, :, 0, :] + (a[:, :, 0, :] * b[:, :, 0, :] - a[:, :, 1, :] * b[:, :, 1, :]) * scaling
    out[:, :, 1, :] = ab[:, :, 1, :] + (a[:, :, 0, :] * b[:, :, 1, :] + a[:, :, 1, :] * b[:, :, 0, :]) * scaling
    if not is_complex:  # DC and Nyquist are two real products (avx:1974-1978)
        out[:, 0, 0, 0] = ab[:, 0, 0, 0] + a[:, 0, 0, 0] * b[:, 0, 0, 0] * scaling
        out[:, 0, 1, 0] = ab[:, 0, 1, 0] + a[:, 0, 1, 0] * b[:, 0, 1, 0] * scaling
    return out.astype(np.float32).reshape(shp)


def np_accumulate(a, b) -> np.ndarray:
    return (np.asarray(a, np.float32) + np.asarray(b, np.float32)).astype(np.float32)


def rel_l2(test, ref) -> float:
    t = np.asarray(test, np.float64).ravel()
    r = np.asarray(ref, np.float64).ravel()
    d = np.linalg.norm(r)
    return float(np.linalg.norm(t - r) / d) if d > 0 else float(np.linalg.norm(t))


def parity_tol(N: int) -> float:
    """north_star: relative L2 error <= 1e-6 * log2(N)."""
    return 1e-6 * np.log2(N)


# reference test signals (test/test.cpp:23-27,82-85,142-148,193-197) -- note 3.14f, not pi
def ref_signal(N: int, is_complex: bool, freq_hz: float = 100.0) -> np.ndarray:
    i = np.arange(N, dtype=np.float32)
    w = np.float32(3.14) * (np.float32(freq_hz) / np.float32(48000.0))
    if is_complex:
        out = np.empty(2 * N, dtype=np.float32)
        out[0::2] = np.sin((w * i).astype(np.float32))
        out[1::2] = np.cos((w * i).astype(np.float32))
        return out
    return np.sin((w * i).astype(np.float32)).astype(np.float32)


# --------------------------------------------------------------------------------------------------
# C restatement
# --------------------------------------------------------------------------------------------------
class COracle:
    def __init__(self, path: str = ORACLE_SO):
        self.lib = C.CDLL(path)
        L = self.lib
        L.oracle_simd_width.argtypes = [C.c_int, C.c_int, C.c_int]
        L.oracle_simd_width.restype = C.c_int
        L.oracle_unordered_map.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]
        L.oracle_transform.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _fp, _fp]
        L.oracle_transform.restype = C.c_int
        L.oracle_convolve.argtypes = [C.c_int, C.c_int, C.c_int, _fp, _fp, _fp, C.c_float]
        L.oracle_accumulate.argtypes = [_fp, _fp, _fp, C.c_int]

    def simd_width(self, N, is_complex, use_avx=True):
        return self.lib.oracle_simd_width(N, int(is_complex), int(use_avx))

    def unordered_map(self, N, is_complex, W):
        m = np.empty(2 * N if is_complex else N, dtype=np.int32)
        self.lib.oracle_unordered_map(N, int(is_complex), W, m.ctypes.data_as(C.POINTER(C.c_int)))
        return m

    def transform(self, x, N, is_complex, W, backward, ordered):
        x = np.ascontiguousarray(x, dtype=np.float32)
        out = np.empty_like(x)
        rc = self.lib.oracle_transform(N, int(is_complex), W, int(backward), int(ordered), _ptr(x), _ptr(out))
        if rc != 0:
            raise ValueError("unsupported N/W")
        return out

    def convolve(self, a, b, ab, N, is_complex, W, scaling):
        a, b = (np.ascontiguousarray(t, dtype=np.float32) for t in (a, b))
        ab = np.array(ab, dtype=np.float32, copy=True)
        self.lib.oracle_convolve(N, int(is_complex), W, _ptr(a), _ptr(b), _ptr(ab), scaling)
        return ab

    def accumulate(self, a, b):
        a, b = (np.ascontiguousarray(t, dtype=np.float32) for t in (a, b))
        ab = np.empty_like(a)
        self.lib.oracle_accumulate(_ptr(a), _ptr(b), _ptr(ab), a.size)
        return ab


# --------------------------------------------------------------------------------------------------
# the unmodified reference
# --------------------------------------------------------------------------------------------------
class RefLib:
    """The reference's own C API (chowdsp_fft.h:81-163) + oracle/ref_driver.cpp batch drivers."""

    def __init__(self, path: str = REF_SO):
        self.lib = C.CDLL(path)
        L = self.lib
        L.fft_new_setup.argtypes = [C.c_int, C.c_int, C.c_bool]
        L.fft_new_setup.restype = C.c_void_p
        L.fft_destroy_setup.argtypes = [C.c_void_p]
        L.fft_simd_width_bytes.argtypes = [C.c_void_p]
        L.fft_simd_width_bytes.restype = C.c_int
        for fn in (L.fft_transform, L.fft_transform_unordered):
            fn.argtypes = [C.c_void_p, _fp, _fp, _fp, C.c_int]
            fn.restype = None
        L.fft_convolve_unordered.argtypes = [C.c_void_p, _fp, _fp, _fp, C.c_float]
        L.fft_convolve_unordered.restype = None
        L.fft_accumulate.argtypes = [C.c_void_p, _fp, _fp, _fp, C.c_int]
        L.fft_accumulate.restype = None
        L.ref_transform_batched.argtypes = [C.c_int] * 5 + [_fp, _fp, C.c_long, C.c_long, C.c_long, C.c_int]
        L.ref_transform_batched.restype = C.c_double
        L.ref_transform_batched_reps.argtypes = [C.c_int] * 5 + [_fp, _fp, C.c_long, C.c_long, C.c_long, C.c_int, C.c_int]
        L.ref_transform_batched_reps.restype = C.c_double
        L.ref_convolve_batched.argtypes = [C.c_int] * 3 + [_fp, _fp, _fp] + [C.c_long] * 4 + [C.c_float, C.c_int]
        L.ref_convolve_batched.restype = C.c_double
        L.ref_partitioned_convolve.argtypes = [C.c_int] * 3 + [_fp] * 4 + [C.c_long, C.c_int, C.c_int, C.c_int]
        L.ref_partitioned_convolve.restype = C.c_double
        L.ref_hardware_threads.restype = C.c_int

    # ---- single calls -------------------------------------------------------------------------
    def new_setup(self, N, is_complex, use_avx=True):
        s = self.lib.fft_new_setup(N, FFT_COMPLEX if is_complex else FFT_REAL, use_avx)
        # reference quirk: unsupported N comes back as (void*)1 on AVX builds
        return None if (s is None or s == 1) else s

    def width(self, setup) -> int:
        return self.lib.fft_simd_width_bytes(C.c_void_p(setup)) // 4

    def hardware_threads(self) -> int:
        return self.lib.ref_hardware_threads()

    def transform(self, x, N, is_complex, backward, ordered, use_avx=True, nthreads=1):
        """x: [..., nfloats]; every row is one reference call. Returns (out, W)."""
        x = np.asarray(x, dtype=np.float32)
        nfl = 2 * N if is_complex else N
        lead = x.shape[:-1]
        xin = aligned_copy(x)
        out = aligned_empty(xin.size)
        batch = xin.size // nfl
        secs = self.lib.ref_transform_batched(N, int(is_complex), int(use_avx), int(backward), int(ordered),
                                              _ptr(xin), _ptr(out), batch, nfl, nfl, nthreads)
        if secs < 0:
            raise ValueError(f"reference rejects N={N}")
        return np.array(out).reshape(*lead, nfl), simd_width(N, is_complex, use_avx)

    def transform_timed(self, xin, out, N, is_complex, backward, ordered, batch, in_stride, out_stride,
                        nthreads, use_avx=True, reps: int = 1) -> float:
        """seconds of `reps` passes over the batch; plan, threads and work buffers are created outside the timed region"""
        return self.lib.ref_transform_batched_reps(N, int(is_complex), int(use_avx), int(backward), int(ordered),
                                                   _ptr(xin), _ptr(out), batch, in_stride, out_stride, nthreads, reps)

    def convolve(self, a, b, ab, N, is_complex, scaling, use_avx=True):
        nfl = 2 * N if is_complex else N
        a, b, ab = (aligned_copy(t) for t in (a, b, ab))
        batch = ab.size // nfl
        secs = self.lib.ref_convolve_batched(N, int(is_complex), int(use_avx), _ptr(a), _ptr(b), _ptr(ab), batch,
                                             nfl if a.size > nfl else 0, nfl if b.size > nfl else 0, nfl, scaling, 1)
        if secs < 0:
            raise ValueError(f"reference rejects N={N}")
        return np.array(ab)

    def accumulate(self, a, b, N, is_complex=False, use_avx=True):
        s = self.new_setup(N, is_complex, use_avx)
        a, b = aligned_copy(a), aligned_copy(b)
        ab = aligned_empty(a.size)
        self.lib.fft_accumulate(C.c_void_p(s), _ptr(a), _ptr(b), _ptr(ab), a.size)
        self.lib.fft_destroy_setup(C.c_void_p(s))
        return np.array(ab)

    def partitioned_convolve(self, x, h, N, P, nthreads=1, use_avx=True):
        """x [channels, blocks*N/2]; h [channels, P, N] unordered spectra -> (y, fdl, secs)."""
        x = np.asarray(x, np.float32)
        channels, total = x.shape
        B = N // 2
        blocks = total // B
        xa, ha = aligned_copy(x), aligned_copy(h)
        fdl = aligned_empty(channels * P * N)
        fdl[:] = 0
        y = aligned_empty(channels * blocks * B)
        secs = self.lib.ref_partitioned_convolve(N, P, int(use_avx), _ptr(xa), _ptr(ha), _ptr(fdl), _ptr(y),
                                                 channels, blocks, 0, nthreads)
        return np.array(y).reshape(channels, blocks * B), np.array(fdl).reshape(channels, P, N), secs


@lru_cache(maxsize=1)
def load_ref() -> RefLib | None:
    return RefLib() if os.path.exists(REF_SO) else None


@lru_cache(maxsize=1)
def load_c() -> COracle:
    if not os.path.exists(ORACLE_SO):
        build()
    return COracle()


def np_partitioned_convolve(x, h, N: int, P: int, W: int, scaling: float | None = None):
    """Uniform partitioned overlap-save convolution restated on top of np_transform / np_convolve, block
    by block, exactly the call sequence of oracle/ref_driver.cpp::ref_partitioned_convolve (itself the
    reference API sequence chowdsp_fft.h:145 ; P x :154 ; :145).  x [channels, blocks*N/2] samples,
    h [channels, P, N] unordered spectra.  Returns (y [channels, blocks*N/2], fdl [channels, P, N])."""
    x = np.asarray(x, np.float32)
    channels, total = x.shape
    B = N // 2
    blocks = total // B
    scaling = 1.0 / N if scaling is None else scaling
    fdl = np.zeros((channels, P, N), np.float32)
    y = np.zeros((channels, blocks * B), np.float32)
    xpad = np.concatenate([np.zeros((channels, B), np.float32), x], axis=1)
    for t in range(blocks):
        win = xpad[:, t * B:t * B + N]
        fdl[:, t % P] = np_transform(win, N, False, W, False, False)
        acc = np.zeros((channels, N), np.float32)
        for p in range(min(t + 1, P)):
            acc = np_convolve(fdl[:, (t - p) % P], h[:, p], acc, N, False, W, scaling)
        y[:, t * B:(t + 1) * B] = np_transform(acc, N, False, W, True, False)[:, B:]
    return y, fdl


def np_istft_overlap_add(spectra, N: int, hop: int, W: int, ordered: bool, window=None, scale: float = 1.0) -> np.ndarray:
    """Overlap-add synthesis restated as the caller-side loop the reference leaves to its users: per frame one
    fft_transform / fft_transform_unordered (BACKWARD) (chowdsp_fft.h:138,145), a window multiply, and a running sum
    of the overlapping parts at hop distance (fft_accumulate, chowdsp_fft.h:160, is the reference's primitive for
    that sum; test/test.cpp:214-218).  spectra [channels, frames, N]; returns [channels, (frames-1)*hop + N]
    float32, sums carried in float64."""
    spectra = np.asarray(spectra, np.float32)
    channels, frames, n = spectra.shape
    assert n == N and 0 < hop <= N
    fr = np_transform(spectra.reshape(-1, N), N, False, W, True, ordered).astype(np.float64).reshape(channels, frames, N) * scale
    if window is not None:
        fr = fr * np.asarray(window, np.float32).astype(np.float64)
    out = np.zeros((channels, (frames - 1) * hop + N), np.float64)
    for f in range(frames):
        out[:, f * hop:f * hop + N] += fr[:, f]
    return out.astype(np.float32)


# --------------------------------------------------------------------------------------------------
# the JUCE adapter's conventions (chowdsp_fft_juce/chowdsp_fft_juce.cpp), restated on top of np_transform
# --------------------------------------------------------------------------------------------------
def np_juce_perform(x, N: int, inverse: bool) -> np.ndarray:
    """ChowDSP_FFT::perform (chowdsp_fft_juce.cpp:32-46): ordered complex transform, inverse scaled by 1/N."""
    y = np_transform(x, N, True, 8, inverse, True)
    return (y.astype(np.float64) / N).astype(np.float32) if inverse else y


def np_juce_real_forward(x, N: int, ignore_negative_freqs: bool) -> np.ndarray:
    """performRealOnlyForwardTransform (chowdsp_fft_juce.cpp:48-66): rows of 2N floats; the first N hold the
    samples on entry; on return bins 0..N/2 as interleaved complex (Nyquist moved from float 1 to float N, the
    imaginary parts of DC and Nyquist zero) and, unless ignored, bin N/2+i = conj(bin N/2-i).  Floats the adapter
    leaves untouched keep their input value."""
    x = np.array(x, np.float32, copy=True)
    rows = x.reshape(-1, x.shape[-1])
    assert rows.shape[1] >= (N + 2 if ignore_negative_freqs else 2 * N)
    packed = np_transform(rows[:, :N], N, False, 8, False, True)
    rows[:, :N] = packed
    rows[:, N] = packed[:, 1]
    rows[:, N + 1] = 0.0
    rows[:, 1] = 0.0
    if not ignore_negative_freqs:
        c = rows[:, 0:N + 2:2] + 1j * rows[:, 1:N + 2:2]          # bins 0..N/2
        mirror = np.conj(c[:, N // 2 - 1:0:-1])                   # bins N/2-1 .. 1 -> N/2+1 .. N-1
        rows[:, N + 2:2 * N:2] = mirror.real
        rows[:, N + 3:2 * N:2] = mirror.imag
    return rows.reshape(x.shape)


def np_juce_real_inverse(x, N: int) -> np.ndarray:
    """performRealOnlyInverseTransform (chowdsp_fft_juce.cpp:68-84): bins 0..N/2 interleaved in, N samples out / N."""
    x = np.array(x, np.float32, copy=True)
    rows = x.reshape(-1, x.shape[-1])
    packed = rows[:, :N].copy()
    packed[:, 1] = rows[:, N]
    rows[:, 1] = rows[:, N]  # the adapter's own in-place fix-up is visible in the buffer only until the transform overwrites it
    rows[:, :N] = (np_transform(packed, N, False, 8, True, True).astype(np.float64) / N).astype(np.float32)
    return rows.reshape(x.shape)
