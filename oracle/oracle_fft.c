/*
 * TEST INFRASTRUCTURE ONLY -- CPU restatement (plain C, double-precision internals) of the semantics of
 * chowdsp_fft's hot path.  It is the checker for the CUDA path; nothing under chowdsp_fft_b200/ may
 * call, link or load it (only tests/, bench.py's cpu_baseline leg and __graft_entry__.smoke()).
 *
 * Parity status: PINNED.  tests/test_oracle.py checks every function here against
 *   (a) the unmodified reference compiled into oracle/_ref/libchowdsp_fft_ref.so (when present), and
 *   (b) tests/golden/*.npz, which were produced by that same reference build (tests/gen_golden.py).
 * The reference ships no golden vectors of its own and its tests compare against pffft, which is not
 * vendored, so (a)/(b) are the pins.
 *
 * What is restated (reference file:line, AVX file = /root/reference/simd/chowdsp_fft_impl_avx.cpp):
 *   oracle_simd_width      size rules + AVX/SSE choice      chowdsp_fft.cpp:258-280, common.hpp:168-177,216-225
 *   oracle_unordered_map   unordered <-> ordered permutation pffft_zreorder avx:1780-1839 (sse:1469-1515)
 *   oracle_transform       ordered / unordered DFT          pffft_transform_internal avx:1848-1935
 *   oracle_convolve        ab += a*b*scaling                pffft_convolve_internal avx:1937-1979
 *   oracle_accumulate      ab = a + b                       fft_accumulate_internal avx:1981-1994
 * The arithmetic inside oracle_transform is NOT FFTPACK's pass structure: it is a textbook radix-2 FFT
 * in double precision, rounded once to fp32.  The parity metric (relative L2 <= 1e-6*log2 N) is
 * insensitive to butterfly order; what must be exact -- formats, signs, scaling, the unordered index
 * map, the DC/Nyquist convention -- is what this file pins down.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

static int is_pow2 (long n) { return n > 0 && (n & (n - 1)) == 0; }
/* sizes the reference's decompose() accepts: only the factors 2, 3 and 5 (common.hpp:51-75) */
static int is_235 (long n)
{
    if (n <= 0)
        return 0;
    while (n % 2 == 0) n /= 2;
    while (n % 3 == 0) n /= 3;
    while (n % 5 == 0) n /= 5;
    return n == 1;
}

/* SIMD width (in floats) the reference would pick for this N: 8 (AVX, handle untagged), 4 (SSE) or
 * 0 = unsupported.  Real needs N % (2*W*W) == 0, complex N % (W*W) == 0 (common.hpp:168-177); AVX is
 * tried first when requested (chowdsp_fft.cpp:262-273).  N = 2^a 3^b 5^c (common.hpp:51-75). */
int oracle_simd_width (int N, int is_complex, int use_avx)
{
    if (! is_235 (N))
        return 0;
    const int w_first = use_avx ? 8 : 4;
    for (int W = w_first; W >= 4; W /= 2)
    {
        const int need = is_complex ? W * W : 2 * W * W;
        if (N % need == 0)
            return W;
    }
    return 0;
}

static void fft_pow2 (double* re, double* im, long n, int sign);

/* DFT of n complex doubles for any n: the radix-2 FFT below for powers of two, else the O(n^2) definition with an
 * exact-index twiddle table (the non-power-of-two sizes of the reference's tests are <= 9216). */
static void dft_any (double* re, double* im, long n, int sign)
{
    if (is_pow2 (n))
    {
        fft_pow2 (re, im, n, sign);
        return;
    }
    double* c = (double*) malloc (sizeof (double) * (size_t) n);
    double* s = (double*) malloc (sizeof (double) * (size_t) n);
    double* xr = (double*) malloc (sizeof (double) * (size_t) n);
    double* xi = (double*) malloc (sizeof (double) * (size_t) n);
    for (long t = 0; t < n; ++t)
    {
        const double ang = sign * 2.0 * M_PI * (double) t / (double) n;
        c[t] = cos (ang);
        s[t] = sin (ang);
        xr[t] = re[t];
        xi[t] = im[t];
    }
    for (long k = 0; k < n; ++k)
    {
        double ar = 0.0, ai = 0.0;
        long t = 0;
        for (long m = 0; m < n; ++m)
        {
            ar += xr[m] * c[t] - xi[m] * s[t];
            ai += xr[m] * s[t] + xi[m] * c[t];
            t += k;
            if (t >= n)
                t -= n;
        }
        re[k] = ar;
        im[k] = ai;
    }
    free (c); free (s); free (xr); free (xi);
}

/* In-place iterative radix-2 DIT FFT on n complex doubles; sign = -1 forward, +1 backward; unscaled. */
static void fft_pow2 (double* re, double* im, long n, int sign)
{
    for (long i = 1, j = 0; i < n; ++i)
    {
        long bit = n >> 1;
        for (; j & bit; bit >>= 1)
            j ^= bit;
        j ^= bit;
        if (i < j)
        {
            double t = re[i]; re[i] = re[j]; re[j] = t;
            t = im[i]; im[i] = im[j]; im[j] = t;
        }
    }
    for (long len = 2; len <= n; len <<= 1)
    {
        const long half = len >> 1;
        for (long k = 0; k < half; ++k)
        {
            const double ang = sign * 2.0 * M_PI * (double) k / (double) len;
            const double wr = cos (ang), wi = sin (ang);
            for (long s = k; s < n; s += len)
            {
                const long o = s + half;
                const double xr = re[o] * wr - im[o] * wi;
                const double xi = re[o] * wi + im[o] * wr;
                re[o] = re[s] - xr; im[o] = im[s] - xi;
                re[s] += xr;        im[s] += xi;
            }
        }
    }
}

/* map[u] = ordered float slot held at unordered float slot u.  nfloats = 2N (complex) or N (real).
 * Ordered formats: complex = interleaved (re,im) in natural bin order; real = pffft packing
 * [X0.re, X_{N/2}.re, X1.re, X1.im, ...] (avx:1780-1839, comment sse:912).  The unordered buffer is a
 * sequence of W-float vectors, vector 2k = W real parts, 2k+1 = the matching imaginary parts. */
void oracle_unordered_map (int N, int is_complex, int W, int* map)
{
    if (is_complex)
    {
        const int L = N / W; /* bins per lane-row */
        for (int k = 0; k < N / W; ++k)
        {
            const int b = k / W, r = k % W;
            for (int j = 0; j < W; ++j)
            {
                const int bin = r * L + b * W + j;
                map[(2 * k) * W + j] = 2 * bin;
                map[(2 * k + 1) * W + j] = 2 * bin + 1;
            }
        }
    }
    else
    {
        const int Q = N / (2 * W);
        for (int k = 0; k < Q; ++k)
        {
            const int b = k / W, r = k % W;
            for (int j = 0; j < W; ++j)
            {
                const int m = b * W + j;
                const int bin = r * Q + ((r & 1) ? (Q - m) % Q : m); /* odd rows are stored reversed */
                map[(2 * k) * W + j] = 2 * bin;
                map[(2 * k + 1) * W + j] = 2 * bin + 1;
            }
        }
    }
}

/* fft_transform (ordered=1) / fft_transform_unordered (ordered=0).  Unscaled in both directions.
 * in/out: N floats (real) or 2N floats (complex); may alias.  Returns 0, or -1 for an unsupported N/W. */
int oracle_transform (int N, int is_complex, int W, int backward, int ordered, const float* in, float* out)
{
    if (! is_235 (N) || (W != 4 && W != 8) || N % (is_complex ? W * W : 2 * W * W) != 0)
        return -1;
    const long nfloats = is_complex ? 2L * N : N;
    double* re = (double*) malloc (sizeof (double) * (size_t) N);
    double* im = (double*) malloc (sizeof (double) * (size_t) N);
    float* freq = (float*) malloc (sizeof (float) * (size_t) nfloats); /* ordered-format staging */
    int* map = NULL;
    if (! ordered)
    {
        map = (int*) malloc (sizeof (int) * (size_t) nfloats);
        oracle_unordered_map (N, is_complex, W, map);
    }

    if (! backward)
    {
        for (long n = 0; n < N; ++n)
        {
            re[n] = is_complex ? in[2 * n] : in[n];
            im[n] = is_complex ? in[2 * n + 1] : 0.0;
        }
        dft_any (re, im, N, -1); /* forward kernel e^{-2 pi i k n / N} */
        if (is_complex)
            for (long k = 0; k < N; ++k) { freq[2 * k] = (float) re[k]; freq[2 * k + 1] = (float) im[k]; }
        else
        {
            freq[0] = (float) re[0];
            freq[1] = (float) re[N / 2]; /* Nyquist rides in the imaginary slot of bin 0 */
            for (long k = 1; k < N / 2; ++k) { freq[2 * k] = (float) re[k]; freq[2 * k + 1] = (float) im[k]; }
        }
        if (ordered)
            memcpy (out, freq, sizeof (float) * (size_t) nfloats);
        else
            for (long u = 0; u < nfloats; ++u)
                out[u] = freq[map[u]];
    }
    else
    {
        if (ordered)
            memcpy (freq, in, sizeof (float) * (size_t) nfloats);
        else
            for (long u = 0; u < nfloats; ++u)
                freq[map[u]] = in[u];
        if (is_complex)
            for (long k = 0; k < N; ++k) { re[k] = freq[2 * k]; im[k] = freq[2 * k + 1]; }
        else
        {
            re[0] = freq[0]; im[0] = 0.0;
            re[N / 2] = freq[1]; im[N / 2] = 0.0;
            for (long k = 1; k < N / 2; ++k)
            {
                re[k] = freq[2 * k];     im[k] = freq[2 * k + 1];
                re[N - k] = freq[2 * k]; im[N - k] = -(double) freq[2 * k + 1]; /* Hermitian extension */
            }
        }
        dft_any (re, im, N, +1); /* unscaled: BACKWARD(FORWARD(x)) = N x (chowdsp_fft.h:128-129) */
        for (long n = 0; n < N; ++n)
        {
            if (is_complex) { out[2 * n] = (float) re[n]; out[2 * n + 1] = (float) im[n]; }
            else out[n] = (float) re[n];
        }
    }
    free (re); free (im); free (freq); free (map);
    return 0;
}

/* fft_convolve_unordered: ab += (a .* b) * scaling on unordered spectra (avx:1937-1979).  Complex
 * products on (vector 2k, vector 2k+1) pairs; for a REAL setup float slots 0 and W are two independent
 * real products (DC and Nyquist).  a, b, ab may alias. */
void oracle_convolve (int N, int is_complex, int W, const float* a, const float* b, float* ab, float scaling)
{
    const long nfloats = is_complex ? 2L * N : N;
    const double ar0 = a[0], ai0 = a[W], br0 = b[0], bi0 = b[W], abr0 = ab[0], abi0 = ab[W];
    for (long v = 0; v < nfloats / (2 * W); ++v)
    {
        const long base = v * 2 * W;
        for (int j = 0; j < W; ++j)
        {
            const double ar = a[base + j], ai = a[base + W + j];
            const double br = b[base + j], bi = b[base + W + j];
            const double pr = ar * br - ai * bi, pi = ar * bi + ai * br;
            ab[base + j] = (float) ((double) ab[base + j] + pr * scaling);
            ab[base + W + j] = (float) ((double) ab[base + W + j] + pi * scaling);
        }
    }
    if (! is_complex)
    {
        ab[0] = (float) (abr0 + ar0 * br0 * scaling);
        ab[W] = (float) (abi0 + ai0 * bi0 * scaling);
    }
}

/* fft_accumulate: ab[i] = a[i] + b[i] for i < n (avx:1981-1994); fp32 add is exact to restate. */
void oracle_accumulate (const float* a, const float* b, float* ab, int n)
{
    for (int i = 0; i < n; ++i)
        ab[i] = a[i] + b[i];
}
