// Mixed-radix transforms N = Q * 2^p with Q in {3, 5, 9, 15} on the register / shared-memory machinery of the power-of-two
// kernels (reference: FFTPACK's passf3 / passf5 / radf3 / radf5 / radb3 / radb5,
// /root/reference/simd/chowdsp_fft_impl_avx.cpp:240-277, :356-428, :755-792, :871-953, :1399-1443, :1535-1629; sizes in
// its tests: 96, 192, 384, 480, 640, 768, 9216, test/test.cpp:279-285).
//
// The generic kernel in mixed_kernels.cuh walks run-time radix lists through two shared-memory buffers (1.0 - 2.3 TB/s).  Here a
// transform of M = Q P complex points (P = 2^LOGP) is decimated in time by the odd factor:
//
//   1. Q sub-transforms of P points, x_q[n1] = z[Q n1 + q], run side by side on Q groups of T = P / 16 threads with the
//      stages of fft_kernel (16 points per thread in registers, radix-16 butterflies, padded shared-memory exchanges, the same
//      twiddle tables);
//   2. their spectra Y_q[k'] are left in the exchange regions in natural order; every thread then owns the output columns
//      k' = t + i (Q T): it multiplies Y_q[k'] by W_M^(q k') -- ONE table value per (thread, q), W_M^(q t), times the compile-time
//      constant W_16^(q i), because Q T = M / 16 -- runs a radix-Q butterfly in registers and stores X[k' + P q''], q'' < Q,
//      with coalesced 64-bit stores.
//
// The strided side is the LOAD (z[Q n1 + q]: 8-byte elements at stride 8 Q): the Q groups of a transform read the same lines,
// so it is served by L1.  Backward transforms are conj (FFT (conj z)); real transforms add the split / merge step of the generic
// kernel (through shared memory) around the M-point complex transform of the packed samples; unordered layouts are applied
// while loading / storing with the closed-form maps of mixed_kernels.cuh.  Sizes with other odd parts (25, 27, 45, ...) keep
// the generic kernel.
#pragma once
#include "mixed_kernels.cuh"

namespace cfb
{
struct MixQArgs
{
    const float* in;
    float* out;
    long long in_stride, out_stride; // floats between consecutive transforms
    int batch;
    int kind;           // Kind
    int W;              // 0 = ordered, 4 / 8 = lanes of the unordered layout
    const float2* tw;   // stage twiddles of the P-point transforms (Geo<LOGP, 16> layout, fill_stage_twiddles)
    const float2* wtab; // W_M^t = exp (-2 pi i t / M), t < M
    const float2* rtab; // real plans: exp (-2 pi i k / (2 M)), k <= M / 2
};

template <int LOGP, int Q>
struct MixQGeo
{
    using G = Geo<LOGP, 16>;
    static constexpr int P = G::M, T = G::T, NT = Q * T, M = Q * P;
    static constexpr int BF = (P + NT - 1) / NT;   // output columns per thread (16 / Q, rounded up)
    static constexpr int REGION = G::SMEM_F2;      // float2 slots per sub-transform: exchange region, then its spectrum
    static constexpr int SLOT_F2 = Q * REGION;
    static constexpr int SLOTS = NT >= 256 ? 1 : 256 / NT; // transforms per CTA
    static constexpr int THREADS = SLOTS * NT;
    static constexpr int SMEM_BYTES = SLOTS * SLOT_F2 * 8;
    static constexpr int REG_THREADS = Q == 15 ? 768 : 1024; // resident threads per SM the register budget is sized for (15 points + 14 twiddles live in the butterfly)
    static constexpr int MIN_BLOCKS = REG_THREADS / THREADS < 1 ? 1 : REG_THREADS / THREADS;
    // element n of an M-point sequence kept in the Q regions: region n / P, offset n % P
    static FFT_HD int flat (int n) { return n + (n >> LOGP) * (REGION - P); }
};

// forward DFT of length 9 = 3 x 3 (n = 3 n1 + n2, k = k1 + 3 k2)
FFT_HD void mx_dft9 (float2* u)
{
    constexpr float c1 = 0.766044443118978035f, s1 = 0.642787609686539327f;  // cos / sin (2 pi / 9)
    constexpr float c2 = 0.173648177666930349f, s2 = 0.984807753012208059f;  // cos / sin (4 pi / 9)
    constexpr float c4 = -0.939692620785908384f, s4 = 0.342020143325668733f; // cos / sin (8 pi / 9)
    float2 t[9];
#pragma unroll
    for (int n2 = 0; n2 < 3; ++n2)
    {
        float2 c[3] = { u[n2], u[3 + n2], u[6 + n2] };
        mx_dft3 (c);
#pragma unroll
        for (int k1 = 0; k1 < 3; ++k1)
            t[3 * k1 + n2] = c[k1];
    }
    t[3 * 1 + 1] = mx_mul (t[3 * 1 + 1], make_float2 (c1, -s1)); // W9^(n2 k1)
    t[3 * 1 + 2] = mx_mul (t[3 * 1 + 2], make_float2 (c2, -s2));
    t[3 * 2 + 1] = mx_mul (t[3 * 2 + 1], make_float2 (c2, -s2));
    t[3 * 2 + 2] = mx_mul (t[3 * 2 + 2], make_float2 (c4, -s4));
#pragma unroll
    for (int k1 = 0; k1 < 3; ++k1)
    {
        float2 c[3] = { t[3 * k1], t[3 * k1 + 1], t[3 * k1 + 2] };
        mx_dft3 (c);
#pragma unroll
        for (int k2 = 0; k2 < 3; ++k2)
            u[k1 + 3 * k2] = c[k2];
    }
}
// forward DFT of length 15 = 3 x 5, prime-factor maps (no twiddles): n = (5 n1 + 3 n2) mod 15, k = (10 k1 + 6 k2) mod 15
FFT_HD void mx_dft15 (float2* u)
{
    float2 t[15];
#pragma unroll
    for (int n2 = 0; n2 < 5; ++n2)
    {
        float2 c[3] = { u[(3 * n2) % 15], u[(5 + 3 * n2) % 15], u[(10 + 3 * n2) % 15] };
        mx_dft3 (c);
#pragma unroll
        for (int k1 = 0; k1 < 3; ++k1)
            t[5 * k1 + n2] = c[k1];
    }
#pragma unroll
    for (int k1 = 0; k1 < 3; ++k1)
    {
        float2 c[5] = { t[5 * k1], t[5 * k1 + 1], t[5 * k1 + 2], t[5 * k1 + 3], t[5 * k1 + 4] };
        mx_dft5 (c);
#pragma unroll
        for (int k2 = 0; k2 < 5; ++k2)
            u[(10 * k1 + 6 * k2) % 15] = c[k2];
    }
}
template <int Q>
FFT_HD void mx_dftq (float2* u)
{
    if constexpr (Q == 3)
        mx_dft3 (u);
    else if constexpr (Q == 5)
        mx_dft5 (u);
    else if constexpr (Q == 9)
        mx_dft9 (u);
    else
        mx_dft15 (u);
}

// MODE 0: ordered complex transforms (both directions) -- the common case gets a kernel of its own, so that the unrolled real / unordered
// prologues and epilogues of MODE 1 (everything else) do not cost it registers (measured: 7.0 vs 6.8 TB/s at N = 768)
template <int LOGP, int Q, int MODE>
FFT_HD void mixq_body (const MixQArgs& a)
{
    constexpr bool GENERAL = MODE != 0;
    using X = MixQGeo<LOGP, Q>;
    using G = typename X::G;
    constexpr int P = X::P, T = X::T, NT = X::NT, M = X::M, R = 16;
    FFT_DYN_SMEM (float2, smem);
    const int tid = (int) threadIdx.x;
    const int slot = tid / NT, lt = tid - slot * NT; // transform of this CTA, thread within the transform
    const int q = lt / T, j = lt - q * T;            // sub-transform, thread within the sub-transform
    const long long x = (long long) blockIdx.x * X::SLOTS + slot;
    const bool active = x < a.batch;
    const long long xc = active ? x : (long long) a.batch - 1; // slots past the end redo the last transform and skip the stores
    const float* __restrict__ in = a.in + xc * a.in_stride;
    float* __restrict__ out = a.out + xc * a.out_stride;
    float2* base = smem + slot * X::SLOT_F2;
    float2* sq = base + q * X::REGION;
    const int kind = GENERAL ? a.kind : (a.kind == C2C_BWD ? C2C_BWD : C2C_FWD), W = GENERAL ? a.W : 0;
    const bool backward = kind == C2C_BWD || (GENERAL && kind == C2R);
    const bool staged_in = GENERAL && (kind == C2R || (kind == C2C_BWD && W != 0));

    // ---- input sequence z (conjugated for the backward kinds): straight from global memory, or staged in the regions ----
    // (M = 16 NT: every thread owns 16 elements n = lt + i NT, or the 8 bin pairs (k, M - k), k = lt + i NT < M / 2; the loops are
    // unrolled so that all global loads of a thread are in flight together)
    if (staged_in)
    {
        if (kind == C2C_BWD)
        {
            // bins 2 p and 2 p + 1 are adjacent lanes of one vector (M / W and W are even): (re, re') and (im, im') are 8-byte pairs
            float2 tr[8], ti[8];
#pragma unroll
            for (int i = 0; i < 8; ++i)
            {
                const int p = mixed_upos_complex (2 * (lt + i * NT), M, W);
                tr[i] = __ldg (reinterpret_cast<const float2*> (in + p));
                ti[i] = __ldg (reinterpret_cast<const float2*> (in + p + W));
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
            {
                const int n = 2 * (lt + i * NT);
                sts2 (base + X::flat (n), make_float2 (tr[i].x, -ti[i].x));
                sts2 (base + X::flat (n + 1), make_float2 (tr[i].y, -ti[i].y));
            }
        }
        else // C2R merge: Z'[k] = (X[k] + X*[M-k]) + i conj(w_k) (X[k] - X*[M-k]), Z'[M-k] = conj ((X[k] + X*[M-k]) - i conj(w_k) (X[k] - X*[M-k])),
        {    // stored conjugated; bin 0 carries (DC, Nyquist) and is paired with bin M / 2 (Z'[M/2] = 2 conj X[M/2])
            float2 xa[8], xm[8], w[8];
#pragma unroll
            for (int i = 0; i < 8; ++i)
            {
                const int k = lt + i * NT;
                const int km = k == 0 ? M / 2 : M - k;
                if (W == 0)
                {
                    xa[i] = __ldg (reinterpret_cast<const float2*> (in) + k);
                    xm[i] = __ldg (reinterpret_cast<const float2*> (in) + km);
                }
                else
                {
                    const int pa = mixed_upos_real (k, M, W), pm = mixed_upos_real (km, M, W);
                    xa[i] = make_float2 (__ldg (in + pa), __ldg (in + pa + W));
                    xm[i] = make_float2 (__ldg (in + pm), __ldg (in + pm + W));
                }
                w[i] = __ldg (a.rtab + k);
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
            {
                const int k = lt + i * NT;
                if (k == 0)
                {
                    sts2 (base, make_float2 (xa[i].x + xa[i].y, xa[i].y - xa[i].x)); // conj Z'[0]
                    sts2 (base + X::flat (M / 2), make_float2 (2.f * xm[i].x, 2.f * xm[i].y));
                }
                else
                {
                    const float2 cm = mx_conj (xm[i]);
                    const float2 e = cadd (xa[i], cm), d = csub (xa[i], cm);
                    const float2 wd = mx_mul (d, mx_conj (w[i]));
                    sts2 (base + X::flat (k), make_float2 (e.x - wd.y, -e.y - wd.x));    // conj (e + i wd)
                    sts2 (base + X::flat (M - k), make_float2 (e.x + wd.y, e.y - wd.x)); // conj Z'[M-k] = e - i wd
                }
            }
        }
        __syncthreads();
    }
    float2 v[R];
    if (staged_in)
    {
#pragma unroll
        for (int m = 0; m < R; ++m)
            v[m] = lds2 (base + X::flat (Q * (j + m * T) + q));
    }
    else
    {
        const float2* __restrict__ in2 = reinterpret_cast<const float2*> (in) + (Q * j + q);
#pragma unroll
        for (int m = 0; m < R; ++m)
        {
            const float2 ld = __ldg (in2 + Q * m * T);
            v[m] = backward ? mx_conj (ld) : ld;
        }
    }

    // ---- Q sub-transforms of P points (forward), CTA-wide barriers ----
    Stages<G, -1, 0, false>::run (v, j, sq, a.tw, staged_in);
    __syncthreads(); // every gather (and every read of a staged input) is done: the regions now receive the spectra
#pragma unroll
    for (int m = 0; m < R; ++m)
        sts2 (sq + j + m * T, v[m]); // Y_q[k'], natural order
    __syncthreads();

    // ---- twiddle + radix-Q butterfly over q for the columns k' = lt + i NT ----
    const bool to_global = ! GENERAL || kind == C2C_BWD || kind == C2R || (kind == C2C_FWD && W == 0);
    float2 wq[Q];
#pragma unroll
    for (int qq = 1; qq < Q; ++qq)
        wq[qq] = __ldg (a.wtab + qq * lt); // W_M^(q lt), q lt < Q NT <= M
    float2* __restrict__ out2 = reinterpret_cast<float2*> (out);
#pragma unroll
    for (int i = 0; i < X::BF; ++i)
    {
        const int k = lt + i * NT;
        if (X::BF * NT == P || k < P)
        {
            float2 u[Q];
#pragma unroll
            for (int qq = 0; qq < Q; ++qq)
                u[qq] = lds2 (base + qq * X::REGION + k);
#pragma unroll
            for (int qq = 1; qq < Q; ++qq)
                u[qq] = cmul_dir<-1> (mul_w32_rt<-1> (u[qq], 2 * ((qq * i) & 15)), wq[qq]); // W_M^(q k) = W_M^(q lt) W_16^(q i)
            mx_dftq<Q> (u);
            if (to_global)
            {
                if (active)
                {
#pragma unroll
                    for (int qq = 0; qq < Q; ++qq)
                        out2[k + P * qq] = backward ? mx_conj (u[qq]) : u[qq];
                }
            }
            else
            {
#pragma unroll
                for (int qq = 0; qq < Q; ++qq)
                    sts2 (base + qq * X::REGION + k, u[qq]); // X[k + P q''] in place (this thread's own slots)
            }
        }
    }
    if (to_global)
        return;
    __syncthreads();

    // ---- spectrum in the regions (natural order): unordered complex store / real split ----
    if (! active)
        return;
    if (kind == C2C_FWD)
    {
#pragma unroll
        for (int i = 0; i < 8; ++i) // bin pairs (2 p, 2 p + 1): adjacent lanes of one vector, stored as 8-byte (re, re') / (im, im') pairs
        {
            const int n = 2 * (lt + i * NT);
            const float2 v0 = lds2 (base + X::flat (n)), v1 = lds2 (base + X::flat (n + 1));
            const int p = mixed_upos_complex (n, M, W);
            *reinterpret_cast<float2*> (out + p) = make_float2 (v0.x, v1.x);
            *reinterpret_cast<float2*> (out + p + W) = make_float2 (v0.y, v1.y);
        }
    }
    else // R2C: X[k] = E - i w_k D, X[M-k] = conj (E + i w_k D), E, D = (Z[k] +- Z*[M-k]) / 2; bin 0 = (DC, Nyquist), paired with bin M / 2
    {
#pragma unroll
        for (int i = 0; i < 8; ++i)
        {
            const int k = lt + i * NT;
            const int km = k == 0 ? M / 2 : M - k;
            const float2 za = lds2 (base + X::flat (k)), zm = lds2 (base + X::flat (km));
            float2 xa, xm;
            if (k == 0)
            {
                xa = make_float2 (za.x + za.y, za.x - za.y); // (DC, Nyquist)
                xm = mx_conj (zm);                           // X[M/2] = conj Z[M/2]
            }
            else
            {
                const float2 w = __ldg (a.rtab + k);
                const float2 cm = mx_conj (zm);
                const float2 e = make_float2 (0.5f * (za.x + cm.x), 0.5f * (za.y + cm.y));
                const float2 d = make_float2 (0.5f * (za.x - cm.x), 0.5f * (za.y - cm.y));
                const float2 wd = mx_mul (d, w);
                xa = make_float2 (e.x + wd.y, e.y - wd.x);
                xm = make_float2 (e.x - wd.y, -e.y - wd.x);
            }
            if (W == 0)
            {
                out2[k] = xa;
                out2[km] = xm;
            }
            else
            {
                const int pa = mixed_upos_real (k, M, W), pm = mixed_upos_real (km, M, W);
                out[pa] = xa.x;
                out[pa + W] = xa.y;
                out[pm] = xm.x;
                out[pm + W] = xm.y;
            }
        }
    }
}

template <int LOGP, int Q, int MODE>
__global__ void __launch_bounds__ (MixQGeo<LOGP, Q>::THREADS, MixQGeo<LOGP, Q>::MIN_BLOCKS) mixq_kernel (const MixQArgs a)
{
    mixq_body<LOGP, Q, MODE> (a);
}

// host: M = Q 2^logP with Q in {3, 5, 9, 15}, 2^logP in 16 .. 4096, M <= kMixedMaxM ?
inline bool mixq_applies (int M, int& logP, int& Q)
{
    if (M <= 0 || M > kMixedMaxM)
        return false;
    int p = 0;
    while ((M & 1) == 0)
    {
        M >>= 1;
        ++p;
    }
    if ((M != 3 && M != 5 && M != 9 && M != 15) || p < 4 || p > 12)
        return false;
    logP = p;
    Q = M;
    return true;
}
} // namespace cfb
