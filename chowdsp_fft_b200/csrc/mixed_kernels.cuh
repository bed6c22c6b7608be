// Mixed-radix transforms: sizes N = 2^a 3^b 5^c that are not powers of two.
//
// The reference reaches these sizes through FFTPACK's radix-3 / radix-5 passes (passf3_ps / passf5_ps, radf3/5,
// radb3/5: /root/reference/simd/chowdsp_fft_impl_avx.cpp:240-277, :356-428, :755-792, :871-953, :1399-1443,
// :1535-1629) and exercises 96, 192, 384, 480, 640, 768 and 9216 in its tests (test/test.cpp:279-285).  They are
// outside the power-of-two north star, so this is ONE generic kernel, written for coverage of the drop-in API
// rather than for the roofline: a CTA owns one transform of M complex points (M = N, or N/2 for real plans) in
// shared memory and runs Stockham autosort passes of radix 16, 4, 2, 3 and 5 between two buffers; the twiddles of
// every pass come from a single table W_M^t (fp64 -> fp32).  Backward transforms use
// IFFT (x) = conj (FFT (conj x)), so only forward butterflies exist.  Real transforms are the M-point complex
// transform of the packed pairs plus the same split / merge step as fft_kernel; ordered (pffft packing) and
// unordered (4- / 8-lane) layouts are applied while loading / storing.
// Shared memory: 16 M bytes, so M <= 12288 (N <= 12288 complex, 24576 real).
#pragma once
#include "fft_kernels.cuh"

namespace cfb
{
constexpr int kMixedMaxStages = 16;
constexpr int kMixedMaxM = 12288;
constexpr int kMixedMaxThreads = 512; // 128 registers per thread: the radix-16 pass keeps 16 points + temporaries in registers

struct MixedArgs
{
    const float* in;
    float* out;
    long long in_stride, out_stride; // floats between consecutive transforms
    int batch;
    int M;                           // complex points per transform
    int nstages;
    int radix[kMixedMaxStages];      // product = M; each 16, 4, 2, 3 or 5
    const float2* wtab;              // W_M^t = exp (-2 pi i t / M), t < M
    const float2* rtab;              // real plans: exp (-2 pi i k / (2 M)), k <= M/2
    int kind;                        // Kind
    int W;                           // 0 = ordered, 4 / 8 = lanes of the unordered layout
    int tg;                          // threads per transform (power of two <= blockDim.x): a CTA runs blockDim.x / tg transforms
};

// launch geometry: about one thread per radix-16 butterfly (the radix-4 / 3 / 5 passes loop), at least a warp and at most
// 1024 threads per transform; small transforms share a CTA of 256 threads
inline void mixed_geometry (int M, int& threads, int& tg)
{
    tg = 32;
    while (tg < kMixedMaxThreads && tg * 16 < M)
        tg *= 2;
    threads = tg < 256 ? 256 : tg;
}

// float offset of the real part of a bin inside the unordered layouts, any N (the imaginary part sits W floats
// later); the power-of-two versions with shifts are unordered_pos_* in fft_kernels.cuh
FFT_HD int mixed_upos_complex (int bin, int N, int W)
{
    const int L = N / W;
    const int r = bin / L, rem = bin - r * L;
    const int b = rem / W, lane = rem - b * W;
    return (b * W + r) * 2 * W + lane;
}
FFT_HD int mixed_upos_real (int bin, int M, int W) // bins 0 .. M-1 of a real transform of 2 M samples
{
    const int Q = M / W;
    const int r = bin / Q, mr = bin - r * Q;
    const int m = (r & 1) ? (mr == 0 ? 0 : Q - mr) : mr;
    const int b = m / W, lane = m - b * W;
    return (b * W + r) * 2 * W + lane;
}

FFT_HD float2 mx_mul (float2 a, float2 w) { return make_float2 (a.x * w.x - a.y * w.y, a.x * w.y + a.y * w.x); }
FFT_HD float2 mx_conj (float2 a) { return make_float2 (a.x, -a.y); }
FFT_HD float2 mx_mi (float2 a) { return make_float2 (a.y, -a.x); } // a * (-i)

// forward DFTs of length 2, 3, 4, 5 on u[0..r)
FFT_HD void mx_dft2 (float2* u)
{
    const float2 a = u[0], b = u[1];
    u[0] = cadd (a, b);
    u[1] = csub (a, b);
}
FFT_HD void mx_dft3 (float2* u)
{
    const float2 t1 = cadd (u[1], u[2]);
    const float2 t2 = make_float2 (u[0].x - 0.5f * t1.x, u[0].y - 0.5f * t1.y);
    const float2 d = csub (u[1], u[2]);
    const float2 t3 = mx_mi (make_float2 (0.866025403784438647f * d.x, 0.866025403784438647f * d.y)); // -i sin(2pi/3) (u1 - u2)
    u[0] = cadd (u[0], t1);
    u[1] = cadd (t2, t3);
    u[2] = csub (t2, t3);
}
FFT_HD void mx_dft4 (float2* u)
{
    const float2 t0 = cadd (u[0], u[2]), t1 = csub (u[0], u[2]), t2 = cadd (u[1], u[3]), t3 = mx_mi (csub (u[1], u[3]));
    u[0] = cadd (t0, t2);
    u[1] = cadd (t1, t3);
    u[2] = csub (t0, t2);
    u[3] = csub (t1, t3);
}
FFT_HD void mx_dft5 (float2* u)
{
    constexpr float c1 = 0.309016994374947424f, c2 = -0.809016994374947424f;  // cos (2 pi / 5), cos (4 pi / 5)
    constexpr float s1 = 0.951056516295153572f, s2 = 0.587785252292473129f;   // sin (2 pi / 5), sin (4 pi / 5)
    const float2 a1 = cadd (u[1], u[4]), a2 = cadd (u[2], u[3]);
    const float2 b1 = csub (u[1], u[4]), b2 = csub (u[2], u[3]);
    const float2 p1 = make_float2 (u[0].x + c1 * a1.x + c2 * a2.x, u[0].y + c1 * a1.y + c2 * a2.y);
    const float2 p2 = make_float2 (u[0].x + c2 * a1.x + c1 * a2.x, u[0].y + c2 * a1.y + c1 * a2.y);
    const float2 q1 = mx_mi (make_float2 (s1 * b1.x + s2 * b2.x, s1 * b1.y + s2 * b2.y)); // -i (s1 b1 + s2 b2)
    const float2 q2 = mx_mi (make_float2 (s2 * b1.x - s1 * b2.x, s2 * b1.y - s1 * b2.y)); // -i (s2 b1 - s1 b2)
    u[0] = cadd (u[0], cadd (a1, a2));
    u[1] = cadd (p1, q1);
    u[4] = csub (p1, q1);
    u[2] = cadd (p2, q2);
    u[3] = csub (p2, q2);
}

// forward DFT of length 16 = 4 x 4 in registers: DFT4 over n1 for every n2 (n = 4 n1 + n2), times W16^(n2 k1), DFT4 over n2;
// output index k = k1 + 4 k2
FFT_HD void mx_dft16 (float2* u)
{
    constexpr float c1 = 0.923879532511286756f, s1 = 0.382683432365089772f, h = 0.707106781186547524f;
    float2 t[16];
#pragma unroll
    for (int n2 = 0; n2 < 4; ++n2)
    {
        float2 c[4] = { u[n2], u[4 + n2], u[8 + n2], u[12 + n2] };
        mx_dft4 (c);
#pragma unroll
        for (int k1 = 0; k1 < 4; ++k1)
            t[4 * k1 + n2] = c[k1];
    }
    // W16^(n2 k1), e = n2 k1 in {1, 2, 3, 4, 6, 9}
    t[4 * 1 + 1] = mx_mul (t[4 * 1 + 1], make_float2 (c1, -s1));
    t[4 * 1 + 2] = mx_mul (t[4 * 1 + 2], make_float2 (h, -h));
    t[4 * 1 + 3] = mx_mul (t[4 * 1 + 3], make_float2 (s1, -c1));
    t[4 * 2 + 1] = mx_mul (t[4 * 2 + 1], make_float2 (h, -h));
    t[4 * 2 + 2] = mx_mi (t[4 * 2 + 2]);
    t[4 * 2 + 3] = mx_mul (t[4 * 2 + 3], make_float2 (-h, -h));
    t[4 * 3 + 1] = mx_mul (t[4 * 3 + 1], make_float2 (s1, -c1));
    t[4 * 3 + 2] = mx_mul (t[4 * 3 + 2], make_float2 (-h, -h));
    t[4 * 3 + 3] = mx_mul (t[4 * 3 + 3], make_float2 (-c1, s1));
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1)
    {
        float2 c[4] = { t[4 * k1], t[4 * k1 + 1], t[4 * k1 + 2], t[4 * k1 + 3] };
        mx_dft4 (c);
#pragma unroll
        for (int k2 = 0; k2 < 4; ++k2)
            u[k1 + 4 * k2] = c[k2];
    }
}

// one Stockham pass of radix RDX: src in natural order of the pass input, dst in autosort order.
//   SRC = 1: src is the transform's input in global memory (first pass of the forward kinds: no staging copy)
//   DST = 1 / 2: dst is the transform's output in global memory, stored as is / conjugated (last pass of the kinds whose
//            output is in natural order); `active` = false skips those stores
template <int RDX, int SRC = 0, int DST = 0>
FFT_HD void mixed_pass (const float2* src, float2* dst, int M, int Ns, const float2* __restrict__ wtab, int tid, int nthreads, bool active = true)
{
    const int cols = M / RDX;
    const int tstep = M / (Ns * RDX); // W_(Ns RDX)^(k q) = W_M^(k q tstep)
    const bool ns_pow2 = (Ns & (Ns - 1)) == 0; // true for every power-of-two pass and the first odd-radix one
    for (int j = tid; j < cols; j += nthreads)
    {
        const int k = ns_pow2 ? (j & (Ns - 1)) : j % Ns;
        float2 u[RDX];
#pragma unroll
        for (int q = 0; q < RDX; ++q)
            u[q] = SRC == 1 ? ldg_stream (src + j + q * cols) : lds2 (src + j + q * cols);
        if (Ns > 1)
        {
#pragma unroll
            for (int q = 1; q < RDX; ++q)
                u[q] = mx_mul (u[q], __ldg (wtab + k * q * tstep)); // k q tstep < Ns RDX tstep = M
        }
        if (RDX == 16)
            mx_dft16 (u);
        else if (RDX == 2)
            mx_dft2 (u);
        else if (RDX == 3)
            mx_dft3 (u);
        else if (RDX == 4)
            mx_dft4 (u);
        else
            mx_dft5 (u);
        const int base = (j - k) * RDX + k;
        if (DST == 0)
        {
#pragma unroll
            for (int q = 0; q < RDX; ++q)
                sts2 (dst + base + q * Ns, u[q]);
        }
        else if (active)
        {
#pragma unroll
            for (int q = 0; q < RDX; ++q)
                dst[base + q * Ns] = DST == 2 ? mx_conj (u[q]) : u[q];
        }
    }
}
template <int SRC, int DST>
FFT_HD void mixed_pass_r (int r, const float2* src, float2* dst, int M, int Ns, const float2* __restrict__ wtab, int tid, int nthreads, bool active = true)
{
    if (r == 16)
        mixed_pass<16, SRC, DST> (src, dst, M, Ns, wtab, tid, nthreads, active);
    else if (r == 4)
        mixed_pass<4, SRC, DST> (src, dst, M, Ns, wtab, tid, nthreads, active);
    else if (r == 2)
        mixed_pass<2, SRC, DST> (src, dst, M, Ns, wtab, tid, nthreads, active);
    else if (r == 3)
        mixed_pass<3, SRC, DST> (src, dst, M, Ns, wtab, tid, nthreads, active);
    else
        mixed_pass<5, SRC, DST> (src, dst, M, Ns, wtab, tid, nthreads, active);
}

FFT_HD void mixed_body (const MixedArgs& a)
{
    FFT_DYN_SMEM (float2, smem);
    const int M = a.M, W = a.W, nthreads = a.tg, tid = (int) threadIdx.x % a.tg, grp = (int) threadIdx.x / a.tg;
    const int per_cta = (int) blockDim.x / a.tg;
    float2* A = smem + (size_t) grp * 2 * M;
    float2* B = A + M;
    const bool backward = a.kind == C2C_BWD || a.kind == C2R;
    // all transforms of a CTA run the same passes, so the CTA barriers below line up; slots past the end of the
    // batch redo the last transform and skip the stores
    for (long long x0 = (long long) blockIdx.x * per_cta; x0 < a.batch; x0 += (long long) gridDim.x * per_cta)
    {
        const bool active = x0 + grp < a.batch;
        const long long x = active ? x0 + grp : a.batch - 1;
        const float* __restrict__ in = a.in + x * a.in_stride;
        float* __restrict__ out = a.out + x * a.out_stride;
        // the first pass of the forward kinds reads global memory itself; the last pass of the kinds whose output is in
        // natural order (ordered complex spectra, time-domain signals) writes it itself (two or more passes)
        const bool src_global = (a.kind == C2C_FWD || a.kind == R2C) && a.nstages >= 2;
        const int dst_global = a.nstages < 2 ? 0 : (a.kind == C2C_FWD && W == 0) ? 1 : backward ? 2 : 0;
        // ---- load: A[n] = stage-0 input (conjugated for the backward kinds) ----
        if (src_global)
        {
        }
        else if (a.kind == C2C_FWD || a.kind == R2C)
        {
            for (int n = tid; n < M; n += nthreads)
                A[n] = reinterpret_cast<const float2*> (in)[n];
        }
        else if (a.kind == C2C_BWD)
        {
            for (int n = tid; n < M; n += nthreads)
            {
                float2 v;
                if (W == 0)
                    v = reinterpret_cast<const float2*> (in)[n];
                else
                {
                    const int p = mixed_upos_complex (n, M, W);
                    v = make_float2 (in[p], in[p + W]);
                }
                A[n] = mx_conj (v);
            }
        }
        else // C2R: Z'[k] = (X[k] + X*[M-k]) + i conj(w_k) (X[k] - X*[M-k]), stored conjugated
        {
            for (int k = tid; k < M; k += nthreads)
            {
                const int km = k == 0 ? 0 : M - k;
                float2 xa, xm;
                if (W == 0)
                {
                    xa = reinterpret_cast<const float2*> (in)[k];
                    xm = reinterpret_cast<const float2*> (in)[km];
                }
                else
                {
                    const int pa = mixed_upos_real (k, M, W), pm = mixed_upos_real (km, M, W);
                    xa = make_float2 (in[pa], in[pa + W]);
                    xm = make_float2 (in[pm], in[pm + W]);
                }
                float2 z;
                if (k == 0)
                    z = make_float2 (xa.x + xa.y, xa.x - xa.y); // slot 0 carries (DC, Nyquist)
                else
                {
                    const int kk = k <= M / 2 ? k : M - k;      // w_(M-k) = -conj (w_k)
                    float2 w = __ldg (a.rtab + kk);
                    if (k > M / 2)
                        w = make_float2 (-w.x, w.y);
                    const float2 cm = mx_conj (xm);
                    const float2 e = cadd (xa, cm), d = csub (xa, cm);
                    const float2 wd = mx_mul (d, mx_conj (w));  // conj(w_k) d
                    z = make_float2 (e.x - wd.y, e.y + wd.x);   // e + i conj(w_k) d
                }
                A[k] = mx_conj (z);
            }
        }
        if (! src_global)
            __syncthreads();
        // ---- Stockham passes ----
        float2* src = A;
        float2* dst = B;
        int Ns = 1;
        for (int s = 0; s < a.nstages; ++s)
        {
            const int r = a.radix[s];
            if (s == 0 && src_global)
                mixed_pass_r<1, 0> (r, reinterpret_cast<const float2*> (in), dst, M, Ns, a.wtab, tid, nthreads);
            else if (s == a.nstages - 1 && dst_global == 1)
                mixed_pass_r<0, 1> (r, src, reinterpret_cast<float2*> (out), M, Ns, a.wtab, tid, nthreads, active);
            else if (s == a.nstages - 1 && dst_global == 2)
                mixed_pass_r<0, 2> (r, src, reinterpret_cast<float2*> (out), M, Ns, a.wtab, tid, nthreads, active);
            else
                mixed_pass_r<0, 0> (r, src, dst, M, Ns, a.wtab, tid, nthreads);
            Ns *= r;
            __syncthreads();
            float2* t = src;
            src = dst;
            dst = t;
        }
        // ---- store (src holds the spectrum in natural order) ----
        if (! active || dst_global != 0)
        {
        }
        else if (a.kind == C2C_FWD)
        {
            for (int n = tid; n < M; n += nthreads)
            {
                const float2 v = src[n];
                if (W == 0)
                    reinterpret_cast<float2*> (out)[n] = v;
                else
                {
                    const int p = mixed_upos_complex (n, M, W);
                    out[p] = v.x;
                    out[p + W] = v.y;
                }
            }
        }
        else if (backward)
        {
            for (int n = tid; n < M; n += nthreads)
                reinterpret_cast<float2*> (out)[n] = mx_conj (src[n]);
        }
        else // R2C: X[k] = E - i w_k D, X[M-k] = conj (E + i w_k D), E, D = (Z[k] +- Z*[M-k]) / 2
        {
            for (int k = tid; k <= M / 2; k += nthreads)
            {
                const float2 za = src[k], zm = src[k == 0 ? 0 : M - k];
                float2 xa, xm;
                if (k == 0)
                {
                    xa = make_float2 (za.x + za.y, za.x - za.y); // (DC, Nyquist)
                    xm = xa;
                }
                else
                {
                    const float2 w = __ldg (a.rtab + k);
                    const float2 cm = mx_conj (zm);
                    const float2 e = make_float2 (0.5f * (za.x + cm.x), 0.5f * (za.y + cm.y));
                    const float2 d = make_float2 (0.5f * (za.x - cm.x), 0.5f * (za.y - cm.y));
                    const float2 wd = mx_mul (d, w);
                    xa = make_float2 (e.x + wd.y, e.y - wd.x);   // E - i w D
                    xm = make_float2 (e.x - wd.y, -e.y - wd.x);  // conj (E + i w D)
                }
                const int km = M - k;
                if (W == 0)
                {
                    reinterpret_cast<float2*> (out)[k] = xa;
                    if (k != 0 && km != k)
                        reinterpret_cast<float2*> (out)[km] = xm;
                }
                else
                {
                    const int pa = mixed_upos_real (k, M, W);
                    out[pa] = xa.x;
                    out[pa + W] = xa.y;
                    if (k != 0 && km != k)
                    {
                        const int pm = mixed_upos_real (km, M, W);
                        out[pm] = xm.x;
                        out[pm + W] = xm.y;
                    }
                }
            }
        }
        __syncthreads(); // the buffers are reused by the next transform of this CTA
    }
}

// (a template only so that the kernel can live in this header, which several translation units include)
template <int UNUSED = 0>
__global__ void __launch_bounds__ (kMixedMaxThreads, 2) mixed_kernel (const MixedArgs a)
{
    mixed_body (a);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
// radix schedule 4, 4, .., 2, 3, .., 5, .. for M = 2^a 3^b 5^c; returns the number of stages, 0 if M has other factors
inline int mixed_factor (int M, int* radix)
{
    int n = 0;
    // two radix-4 steps in registers: half the shared-memory round trips (not for tiny transforms, whose radix-16 pass would
    // keep a handful of threads busy)
    const bool use16 = M >= 256;
    while (use16 && M % 16 == 0 && n < kMixedMaxStages)
    {
        radix[n++] = 16;
        M /= 16;
    }
    while (M % 4 == 0 && n < kMixedMaxStages)
    {
        radix[n++] = 4;
        M /= 4;
    }
    for (int r : { 2, 3, 5 })
        while (M % r == 0 && n < kMixedMaxStages)
        {
            radix[n++] = r;
            M /= r;
        }
    return M == 1 ? n : 0;
}
inline void fill_mixed_twiddles (float2* wtab, int M)
{
    const long double two_pi = 2.0L * 3.141592653589793238462643383279502884L;
    for (int t = 0; t < M; ++t)
    {
        const long double ang = -two_pi * (long double) t / (long double) M;
        wtab[t] = make_float2 ((float) cosl (ang), (float) sinl (ang));
    }
}
inline void fill_mixed_real_twiddles (float2* rtab, int M) // M/2 + 1 entries exp (-2 pi i k / (2 M))
{
    const long double pi = 3.141592653589793238462643383279502884L;
    for (int k = 0; k <= M / 2; ++k)
    {
        const long double ang = -pi * (long double) k / (long double) M;
        rtab[k] = make_float2 ((float) cosl (ang), (float) sinl (ang));
    }
}
} // namespace cfb
