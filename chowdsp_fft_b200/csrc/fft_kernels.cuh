// Single-kernel shared-memory Stockham FFT for sm_100a: batched power-of-2 fp32 C2C / R2C / C2R,
// ordered and "unordered" (the reference's SIMD-blocked layout), M = 2^4 .. 2^14 complex points per CTA.
//
// This replaces, on the GPU, the reference's whole per-transform pipeline
//   uninterleave -> cfftf1_ps/rfftf1_ps passes -> pffft_cplx_finalize / pffft_real_finalize -> pffft_zreorder
//   (/root/reference/simd/chowdsp_fft_impl_avx.cpp:1848-1935 and the passes at :206-490, :719-1013,
//    :1016-1351, :1354-1689, :1693-1839)
// with ONE launch that reads each input element from HBM once and writes each output element once.
// It is not a translation of the FFTPACK passes: each thread keeps R (=16) complex points in registers,
// runs radix-16 butterflies, and exchanges data between stages through padded shared memory
// (Stockham autosort, so no bit reversal pass); the real transforms are an N/2-point complex FFT with
// the split/merge step fused in, and the output permutation (natural order or the reference's
// unordered layout) is applied in the epilogue's addressing.
//
// The same source compiles for the host with -DCHOWDSP_EMU (tests/emu): test infrastructure that
// checks index maps / bank conflicts without a GPU.  The product only ever builds the nvcc path.
#pragma once

#ifdef CHOWDSP_EMU
#include "cuda_emu.h"
#define FFT_HD inline
#define FFT_CX constexpr
#define FFT_DYN_SMEM(type, name) type* name = reinterpret_cast<type*> (emu::ctx.smem)
#else
#include <cuda_runtime.h>
#define FFT_HD __device__ __forceinline__
#define FFT_CX __host__ __device__ constexpr
#define FFT_DYN_SMEM(type, name)                   \
    extern __shared__ __align__ (16) char name##_raw[]; \
    type* name = reinterpret_cast<type*> (name##_raw)
#endif

namespace cfb
{
// ---------------------------------------------------------------------------------------------
// transform kinds / kernel arguments
// ---------------------------------------------------------------------------------------------
enum Kind : int
{
    C2C_FWD = 0, // interleaved complex in  -> spectrum out
    C2C_BWD = 1, // spectrum in             -> interleaved complex out
    R2C = 2,     // N real in               -> half spectrum out (N floats)
    C2R = 3      // half spectrum in        -> N real out
};

struct FftArgs
{
    const float* in;
    float* out;
    // transform x = outer * inner + i   reads  in + outer * in_outer + i * in_inner   (floats);
    // plain batches use inner = batch, outer stride 0.  The two-level form is the STFT frame gather.
    long long in_inner, in_outer, out_inner, out_outer;
    int inner;
    int batch;           // total transforms
    int logW;            // unordered layout: 3 = W8 (AVX handle), 2 = W4 (SSE handle)
    const float2* tw;    // stage twiddles, see twiddle_table_len()
    const float2* rtw;   // real split twiddles exp(-2 pi i k / N), k < N/4 (R2C / C2R only)
};

// ---------------------------------------------------------------------------------------------
// geometry
// ---------------------------------------------------------------------------------------------
FFT_CX int ilog2 (int v) { return v <= 1 ? 0 : 1 + ilog2 (v >> 1); }
FFT_CX int ipow (int b, int e) { return e == 0 ? 1 : b * ipow (b, e - 1); }

template <int LOGM_, int R_>
struct Geo
{
    static constexpr int LOGM = LOGM_;
    static constexpr int M = 1 << LOGM_;                 // complex points per transform
    static constexpr int R = R_;                         // complex points held per thread
    static constexpr int LOGR = ilog2 (R_);
    static constexpr int T = M / R;                      // threads per transform
    static constexpr int S = (LOGM + LOGR - 1) / LOGR;   // stages
    static constexpr int RLAST = 1 << (LOGM - (S - 1) * LOGR);
    static constexpr int SMEM_F2 = M + (M >> 4);         // padded float2 slots per transform
    static constexpr int SMEM_F2_UNORD = M + (M >> 3);   // staging image of the unordered layout (W=4 pad is the larger one)
    static_assert (M >= R, "transform smaller than the per-thread radix");
    static FFT_CX int radix (int s) { return s == S - 1 ? RLAST : R; }
    static FFT_CX int ns (int s) { return ipow (R, s); } // product of the radices before stage s
    // float2 offset of stage s's twiddle table (stage 0 has none): [(t-1) * Ns + k], t = 1..r-1
    static FFT_CX int tw_off (int s) { return s <= 1 ? 0 : tw_off (s - 1) + (radix (s - 1) - 1) * ns (s - 1); }
    static constexpr int TW_LEN = S == 1 ? 0 : tw_off (S - 1) + (RLAST - 1) * ns (S - 1);
};

// one padding slot per 16 float2: stride-2^a accesses (a <= 4) and unit-stride accesses are both
// conflict free, and every offset used below folds into an immediate.
FFT_CX int pad (int i) { return i + (i >> 4); }

// ---------------------------------------------------------------------------------------------
// memory helpers (shared-memory ones are instrumented in the emulator)
// ---------------------------------------------------------------------------------------------
FFT_HD float2 lds2 (const float2* p)
{
#ifdef CHOWDSP_EMU
    if (emu::ctx.log)
        emu::ctx.log->push_back ({ (uint32_t) ((const char*) p - emu::ctx.smem), 8, 0 });
#endif
    return *p;
}
FFT_HD void sts2 (float2* p, float2 v)
{
#ifdef CHOWDSP_EMU
    if (emu::ctx.log)
        emu::ctx.log->push_back ({ (uint32_t) ((const char*) p - emu::ctx.smem), 8, 1 });
#endif
    *p = v;
}
FFT_HD void smem_skip() // keeps the per-warp op sequences aligned when a lane is predicated off
{
#ifdef CHOWDSP_EMU
    if (emu::ctx.log)
        emu::ctx.log->push_back ({ 0, 0, 0 });
#endif
}

// ---------------------------------------------------------------------------------------------
// complex arithmetic.  DIR = -1 forward (e^{-i..}), +1 backward.
// ---------------------------------------------------------------------------------------------
// Blackwell packed fp32x2 math (SASS FADD2 / FMUL2 / FFMA2): one instruction per COMPLEX add, two per
// complex multiply.  The operand modifiers of those instructions (half swap, per-half negate, scalar
// broadcast) absorb the (y, x) / (-y, x) shuffles written below, so multiplying by +-i is free.
#ifdef CHOWDSP_EMU
FFT_HD float2 f2_add (float2 a, float2 b) { return make_float2 (a.x + b.x, a.y + b.y); }
FFT_HD float2 f2_sub (float2 a, float2 b) { return make_float2 (a.x - b.x, a.y - b.y); }
FFT_HD float2 f2_mul (float2 a, float2 b) { return make_float2 (a.x * b.x, a.y * b.y); }
FFT_HD float2 f2_fma (float2 a, float2 b, float2 c) { return make_float2 (std::fmaf (a.x, b.x, c.x), std::fmaf (a.y, b.y, c.y)); }
#else
FFT_HD unsigned long long f2_bits (float2 v) { return *reinterpret_cast<unsigned long long*> (&v); }
FFT_HD float2 f2_from (unsigned long long b) { return *reinterpret_cast<float2*> (&b); }
FFT_HD float2 f2_add (float2 a, float2 b)
{
    unsigned long long r;
    asm ("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(f2_bits (a)), "l"(f2_bits (b)));
    return f2_from (r);
}
FFT_HD float2 f2_sub (float2 a, float2 b)
{
    unsigned long long r;
    asm ("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(f2_bits (a)), "l"(f2_bits (b)));
    return f2_from (r);
}
FFT_HD float2 f2_mul (float2 a, float2 b)
{
    unsigned long long r;
    asm ("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(f2_bits (a)), "l"(f2_bits (b)));
    return f2_from (r);
}
FFT_HD float2 f2_fma (float2 a, float2 b, float2 c)
{
    unsigned long long r;
    asm ("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(f2_bits (a)), "l"(f2_bits (b)), "l"(f2_bits (c)));
    return f2_from (r);
}
#endif

FFT_HD float2 cadd (float2 a, float2 b) { return f2_add (a, b); }
FFT_HD float2 csub (float2 a, float2 b) { return f2_sub (a, b); }
// a * w (DIR < 0) or a * conj(w) (DIR > 0); tables always hold the forward twiddle
template <int DIR>
FFT_HD float2 cmul_dir (float2 a, float2 w)
{
    const float2 t = f2_mul (a, make_float2 (w.x, w.x));
    if (DIR < 0)
        return f2_fma (make_float2 (a.y, a.x), make_float2 (-w.y, w.y), t); // (ax wx - ay wy, ay wx + ax wy)
    return f2_fma (make_float2 (a.y, a.x), make_float2 (w.y, -w.y), t);     // (ax wx + ay wy, ay wx - ax wy)
}
// a * (-i) forward, a * (+i) backward
template <int DIR>
FFT_HD float2 mul_mi (float2 a)
{
    return DIR < 0 ? make_float2 (a.y, -a.x) : make_float2 (-a.y, a.x);
}
// a * exp(DIR * 2 pi i * NUM / 16) with compile-time constants
template <int DIR, int NUM>
FFT_HD float2 mul_w16 (float2 a)
{
    constexpr float C1 = 0.923879532511286756f, S1 = 0.382683432365089772f, H = 0.707106781186547524f;
    constexpr int n = NUM & 15;
    if (n == 0)
        return a;
    if (n == 4)
        return mul_mi<DIR> (a);
    if (n == 8)
        return make_float2 (-a.x, -a.y);
    if (n == 12)
        return mul_mi<-DIR> (a);
    // forward twiddle c - i s  with  c = cos(2 pi n/16), s = sin(2 pi n/16)
    constexpr float c = (n == 1 || n == 15) ? C1 : (n == 2 || n == 14) ? H : (n == 3 || n == 13) ? S1
                      : (n == 5 || n == 11) ? -S1 : (n == 6 || n == 10) ? -H : -C1;
    constexpr float sabs = (n == 1 || n == 7 || n == 9 || n == 15) ? S1 : (n == 2 || n == 6 || n == 10 || n == 14) ? H : C1;
    constexpr float s = n < 8 ? sabs : -sabs;
    return cmul_dir<DIR> (a, make_float2 (c, -s));
}

// ---------------------------------------------------------------------------------------------
// register butterflies: in place, natural-order output, elements v[0], v[ST], v[2 ST], ...
// ---------------------------------------------------------------------------------------------
template <int DIR>
FFT_HD void bfly4 (float2& a0, float2& a1, float2& a2, float2& a3)
{
    const float2 t0 = cadd (a0, a2), t1 = csub (a0, a2), t2 = cadd (a1, a3), t3 = mul_mi<DIR> (csub (a1, a3));
    a0 = cadd (t0, t2);
    a1 = cadd (t1, t3);
    a2 = csub (t0, t2);
    a3 = csub (t1, t3);
}

template <int RADIX, int DIR, int ST>
struct RegFft;

template <int DIR, int ST>
struct RegFft<1, DIR, ST>
{
    static FFT_HD void run (float2*) {}
};
template <int DIR, int ST>
struct RegFft<2, DIR, ST>
{
    static FFT_HD void run (float2* v)
    {
        const float2 a = v[0], b = v[ST];
        v[0] = cadd (a, b);
        v[ST] = csub (a, b);
    }
};
template <int DIR, int ST>
struct RegFft<4, DIR, ST>
{
    static FFT_HD void run (float2* v) { bfly4<DIR> (v[0], v[ST], v[2 * ST], v[3 * ST]); }
};
template <int DIR, int ST>
struct RegFft<8, DIR, ST>
{
    static FFT_HD void run (float2* v)
    {
        float2 e0 = v[0], e1 = v[2 * ST], e2 = v[4 * ST], e3 = v[6 * ST];
        float2 o0 = v[ST], o1 = v[3 * ST], o2 = v[5 * ST], o3 = v[7 * ST];
        bfly4<DIR> (e0, e1, e2, e3);
        bfly4<DIR> (o0, o1, o2, o3);
        o1 = mul_w16<DIR, 2> (o1);
        o2 = mul_mi<DIR> (o2);
        o3 = mul_w16<DIR, 6> (o3);
        v[0] = cadd (e0, o0);      v[4 * ST] = csub (e0, o0);
        v[ST] = cadd (e1, o1);     v[5 * ST] = csub (e1, o1);
        v[2 * ST] = cadd (e2, o2); v[6 * ST] = csub (e2, o2);
        v[3 * ST] = cadd (e3, o3); v[7 * ST] = csub (e3, o3);
    }
};
template <int DIR, int ST>
struct RegFft<16, DIR, ST>
{
    // 16 = 4 x 4:  X[k1 + 4 k2] = sum_n2 W16^(n2 k1) W4^(n2 k2) [ sum_n1 a[4 n1 + n2] W4^(n1 k1) ]
    static FFT_HD void run (float2* v)
    {
        float2 b[4][4]; // b[n2][k1]
#pragma unroll
        for (int n2 = 0; n2 < 4; ++n2)
        {
            b[n2][0] = v[(n2) *ST];
            b[n2][1] = v[(n2 + 4) * ST];
            b[n2][2] = v[(n2 + 8) * ST];
            b[n2][3] = v[(n2 + 12) * ST];
            bfly4<DIR> (b[n2][0], b[n2][1], b[n2][2], b[n2][3]);
        }
        b[1][1] = mul_w16<DIR, 1> (b[1][1]);
        b[1][2] = mul_w16<DIR, 2> (b[1][2]);
        b[1][3] = mul_w16<DIR, 3> (b[1][3]);
        b[2][1] = mul_w16<DIR, 2> (b[2][1]);
        b[2][2] = mul_w16<DIR, 4> (b[2][2]);
        b[2][3] = mul_w16<DIR, 6> (b[2][3]);
        b[3][1] = mul_w16<DIR, 3> (b[3][1]);
        b[3][2] = mul_w16<DIR, 6> (b[3][2]);
        b[3][3] = mul_w16<DIR, 9> (b[3][3]);
#pragma unroll
        for (int k1 = 0; k1 < 4; ++k1)
        {
            bfly4<DIR> (b[0][k1], b[1][k1], b[2][k1], b[3][k1]);
            v[(k1) *ST] = b[0][k1];
            v[(k1 + 4) * ST] = b[1][k1];
            v[(k1 + 8) * ST] = b[2][k1];
            v[(k1 + 12) * ST] = b[3][k1];
        }
    }
};

// ---------------------------------------------------------------------------------------------
// position of a bin inside the reference's unordered layouts (pffft_zreorder,
// /root/reference/simd/chowdsp_fft_impl_avx.cpp:1780-1839).  Returns the float offset of the real
// part; the imaginary part sits W floats later.
// ---------------------------------------------------------------------------------------------
template <int LOGN> // complex transform of N = 2^LOGN bins
FFT_HD int unordered_pos_complex (int bin, int logW)
{
    const int logL = LOGN - logW;                   // bins per lane row
    const int r = bin >> logL, rem = bin & ((1 << logL) - 1);
    const int b = rem >> logW, lane = rem & ((1 << logW) - 1);
    return ((((b << logW) + r) * 2) << logW) + lane;
}
template <int LOGM> // real transform of 2M samples: bins 0..M-1 (bin 0 carries DC.re and Nyquist.re)
FFT_HD int unordered_pos_real (int bin, int logW)
{
    const int logQ = LOGM - logW, Q = 1 << logQ;
    const int r = bin >> logQ, mr = bin & (Q - 1);
    const int m = (r & 1) ? ((Q - mr) & (Q - 1)) : mr; // odd rows are stored reversed
    const int b = m >> logW, lane = m & ((1 << logW) - 1);
    return ((((b << logW) + r) * 2) << logW) + lane;
}

// ---------------------------------------------------------------------------------------------
// one Stockham stage on the thread's R registers.  On entry v[m] = x_s[j + m T] (x_s = stage input in
// natural order); on exit of the last stage v[m] = X[j + m T].
// ---------------------------------------------------------------------------------------------
template <class G, int DIR, int STAGE>
FFT_HD void stage_compute (float2 (&v)[G::R], int j, const float2* __restrict__ tw)
{
    constexpr int r = G::radix (STAGE), Ns = G::ns (STAGE), SUB = G::R / r;
    if constexpr (STAGE > 0)
    {
        const float2* __restrict__ t = tw + G::tw_off (STAGE);
#pragma unroll
        for (int u = 0; u < SUB; ++u)
        {
            const int k = (j + u * G::T) & (Ns - 1);
#pragma unroll
            for (int q = 1; q < r; ++q)
                v[u + q * SUB] = cmul_dir<DIR> (v[u + q * SUB], __ldg (t + (q - 1) * Ns + k));
        }
    }
#pragma unroll
    for (int u = 0; u < SUB; ++u)
        RegFft<r, DIR, SUB>::run (&v[u]);
}

// scatter the stage's outputs to shared memory in the order the next stage reads them
template <class G, int STAGE>
FFT_HD void stage_scatter (const float2 (&v)[G::R], int j, float2* s)
{
    constexpr int r = G::radix (STAGE), Ns = G::ns (STAGE), SUB = G::R / r;
#pragma unroll
    for (int u = 0; u < SUB; ++u)
    {
        const int jv = j + u * G::T;
        const int k = jv & (Ns - 1);
        const int base = (jv - k) * r + k;
#pragma unroll
        for (int q = 0; q < r; ++q)
            sts2 (s + pad (base + q * Ns), v[u + q * SUB]);
    }
}

template <class G>
FFT_HD void gather_natural (float2 (&v)[G::R], int j, const float2* s)
{
#pragma unroll
    for (int m = 0; m < G::R; ++m)
        v[m] = lds2 (s + pad (j + m * G::T));
}
template <class G>
FFT_HD void scatter_natural (const float2 (&v)[G::R], int j, float2* s)
{
#pragma unroll
    for (int m = 0; m < G::R; ++m)
        sts2 (s + pad (j + m * G::T), v[m]);
}

template <class G, int DIR, int STAGE>
struct Stages
{
    // from_smem: v was gathered from shared memory just before (a barrier is needed before overwriting it)
    static FFT_HD void run (float2 (&v)[G::R], int j, float2* s, const float2* __restrict__ tw, bool from_smem)
    {
        stage_compute<G, DIR, STAGE> (v, j, tw);
        if constexpr (STAGE < G::S - 1)
        {
            if (STAGE > 0 || from_smem)
                __syncthreads(); // every thread has finished reading the previous exchange
            stage_scatter<G, STAGE> (v, j, s);
            __syncthreads();
            gather_natural<G> (v, j, s);
            Stages<G, DIR, STAGE + 1>::run (v, j, s, tw, true);
        }
    }
};

// ---------------------------------------------------------------------------------------------
// unordered I/O goes through a shared-memory staging image of the spectrum IN THE UNORDERED LAYOUT, so
// that global memory only ever sees linear 128-bit accesses; the permutation is paid on-chip.
// Staging pad: one W-float gap per 2 W^2 floats keeps both the scattered 4-byte accesses (runs of W
// lanes at stride 2 W^2) and the linear 16-byte accesses conflict free.
// ---------------------------------------------------------------------------------------------
FFT_HD int upad (int p, int logW) { return p + ((p >> (2 * logW + 1)) << logW); }

FFT_HD float lds1 (const float* p)
{
#ifdef CHOWDSP_EMU
    if (emu::ctx.log)
        emu::ctx.log->push_back ({ (uint32_t) ((const char*) p - emu::ctx.smem), 4, 0 });
#endif
    return *p;
}
FFT_HD void sts1 (float* p, float v)
{
#ifdef CHOWDSP_EMU
    if (emu::ctx.log)
        emu::ctx.log->push_back ({ (uint32_t) ((const char*) p - emu::ctx.smem), 4, 1 });
#endif
    *p = v;
}
FFT_HD float4 lds4 (const float* p)
{
#ifdef CHOWDSP_EMU
    if (emu::ctx.log)
        emu::ctx.log->push_back ({ (uint32_t) ((const char*) p - emu::ctx.smem), 16, 0 });
#endif
    return *reinterpret_cast<const float4*> (p);
}
FFT_HD void sts4 (float* p, float4 v)
{
#ifdef CHOWDSP_EMU
    if (emu::ctx.log)
        emu::ctx.log->push_back ({ (uint32_t) ((const char*) p - emu::ctx.smem), 16, 1 });
#endif
    *reinterpret_cast<float4*> (p) = v;
}

template <bool REAL_LAYOUT, int LOGM>
FFT_HD float2 staged_load_bin (const float* sf, int bin, int logW)
{
    const int p = REAL_LAYOUT ? unordered_pos_real<LOGM> (bin, logW) : unordered_pos_complex<LOGM> (bin, logW);
    return make_float2 (lds1 (sf + upad (p, logW)), lds1 (sf + upad (p + (1 << logW), logW)));
}
template <bool REAL_LAYOUT, int LOGM>
FFT_HD void staged_store_bin (float* sf, int bin, int logW, float2 v)
{
    const int p = REAL_LAYOUT ? unordered_pos_real<LOGM> (bin, logW) : unordered_pos_complex<LOGM> (bin, logW);
    sts1 (sf + upad (p, logW), v.x);
    sts1 (sf + upad (p + (1 << logW), logW), v.y);
}
// linear 128-bit copies between global memory and the staging image (2M floats = T * R/2 float4)
template <class G>
FFT_HD void staging_fill (float* sf, const float* __restrict__ in, int j, int logW, bool active)
{
#pragma unroll
    for (int i = 0; i < G::R / 2; ++i)
    {
        const int q = j + i * G::T;
        const float4 val = active ? __ldg (reinterpret_cast<const float4*> (in) + q) : make_float4 (0.f, 0.f, 0.f, 0.f);
        sts4 (sf + upad (4 * q, logW), val);
    }
}
template <class G>
FFT_HD void staging_drain (const float* sf, float* __restrict__ out, int j, int logW, bool active)
{
#pragma unroll
    for (int i = 0; i < G::R / 2; ++i)
    {
        const int q = j + i * G::T;
        const float4 val = lds4 (sf + upad (4 * q, logW));
        if (active)
            reinterpret_cast<float4*> (out)[q] = val;
    }
}

// ---------------------------------------------------------------------------------------------
// the kernel.  blockDim.x = T * (transforms per CTA); dynamic smem = transforms per CTA * SMEM_F2 * 8.
// ---------------------------------------------------------------------------------------------
template <int LOGM, int R, int KIND, bool UNORD>
FFT_HD void fft_body (const FftArgs& a)
{
    using G = Geo<LOGM, R>;
    constexpr int DIR = (KIND == C2C_FWD || KIND == R2C) ? -1 : +1;
    constexpr int M = G::M, T = G::T;
    constexpr int SMEM_F2 = UNORD ? G::SMEM_F2_UNORD : G::SMEM_F2;
    FFT_DYN_SMEM (float2, smem);

    const int tid = (int) threadIdx.x;
    const int j = tid & (T - 1);
    const int lt = tid / T;
    const int per_cta = (int) blockDim.x / T;
    const long long x = (long long) blockIdx.x * per_cta + lt;
    const bool active = x < a.batch;
    float2* s = smem + lt * SMEM_F2;
    float* sf = reinterpret_cast<float*> (s); // the same buffer seen as the unordered staging image
    const int logW = a.logW;
    const float2 zero2 = make_float2 (0.f, 0.f);

    const long long xo = active ? x / a.inner : 0, xi = active ? x - xo * a.inner : 0;
    const float* __restrict__ in = a.in + xo * a.in_outer + xi * a.in_inner;
    float* __restrict__ out = a.out + xo * a.out_outer + xi * a.out_inner;

    float2 v[R];
    bool smem_was_read = false; // a barrier is needed before the exchange buffer is overwritten

    // ---- prologue: v[m] = stage-0 input element j + m T -------------------------------------------
    if constexpr (KIND == C2C_FWD || KIND == R2C)
    {
        // interleaved complex, or real samples read as (x[2n], x[2n+1]) pairs
#pragma unroll
        for (int m = 0; m < R; ++m)
            v[m] = active ? __ldg (reinterpret_cast<const float2*> (in) + j + m * T) : zero2;
    }
    else if constexpr (KIND == C2C_BWD)
    {
        if constexpr (UNORD)
        {
            staging_fill<G> (sf, in, j, logW, active);
            __syncthreads();
#pragma unroll
            for (int m = 0; m < R; ++m)
                v[m] = staged_load_bin<false, LOGM> (sf, j + m * T, logW);
            smem_was_read = true;
        }
        else
        {
#pragma unroll
            for (int m = 0; m < R; ++m)
                v[m] = active ? __ldg (reinterpret_cast<const float2*> (in) + j + m * T) : zero2;
        }
    }
    else // C2R: merge step  Z'[k] = (X[k] + X*[M-k]) + i conj(w_k) (X[k] - X*[M-k]),  w_k = e^{-2 pi i k / 2M}
    {
        // fetch this thread's bin pairs (k, M-k), k = j + m T < M/2, into v[2m], v[2m+1];
        // thread 0's first pair is (bin 0 = (DC, Nyquist), bin M/2) instead
        if constexpr (UNORD)
        {
            staging_fill<G> (sf, in, j, logW, active);
            __syncthreads();
        }
#pragma unroll
        for (int m = 0; m < R / 2; ++m)
        {
            const int k = j + m * T;
            const int ka = k, kb = (m == 0 && j == 0) ? M / 2 : M - k;
            if constexpr (UNORD)
            {
                v[2 * m] = staged_load_bin<true, LOGM> (sf, ka, logW);
                v[2 * m + 1] = staged_load_bin<true, LOGM> (sf, kb, logW);
            }
            else
            {
                v[2 * m] = active ? __ldg (reinterpret_cast<const float2*> (in) + ka) : zero2;
                v[2 * m + 1] = active ? __ldg (reinterpret_cast<const float2*> (in) + kb) : zero2;
            }
        }
        if constexpr (UNORD)
            __syncthreads(); // staging image fully consumed before the natural-order image overwrites it
#pragma unroll
        for (int m = 0; m < R / 2; ++m)
        {
            const int k = j + m * T;
            const float2 xa = v[2 * m], xb = v[2 * m + 1];
            if (m == 0 && j == 0)
            {
                sts2 (s + pad (0), make_float2 (xa.x + xa.y, xa.x - xa.y));
                sts2 (s + pad (M / 2), make_float2 (2.f * xb.x, -2.f * xb.y));
            }
            else
            {
                const float2 w = __ldg (a.rtw + k);
                const float2 e = make_float2 (xa.x + xb.x, xa.y - xb.y);
                const float2 d = make_float2 (xa.x - xb.x, xa.y + xb.y);
                const float2 wd = cmul_dir<+1> (d, w); // conj(w) * d
                sts2 (s + pad (k), make_float2 (e.x - wd.y, e.y + wd.x));
                sts2 (s + pad (M - k), make_float2 (e.x + wd.y, wd.x - e.y));
            }
        }
        __syncthreads();
        gather_natural<G> (v, j, s);
        smem_was_read = true;
    }

    // ---- the stages -----------------------------------------------------------------------------
    Stages<G, DIR, 0>::run (v, j, s, a.tw, smem_was_read);
    if (G::S > 1)
        smem_was_read = true;

    // ---- epilogue -------------------------------------------------------------------------------
    if constexpr (KIND == C2C_BWD || KIND == C2R)
    {
        if (active)
        {
#pragma unroll
            for (int m = 0; m < R; ++m)
                reinterpret_cast<float2*> (out)[j + m * T] = v[m];
        }
    }
    else if constexpr (KIND == C2C_FWD)
    {
        if constexpr (UNORD)
        {
            if (smem_was_read)
                __syncthreads();
#pragma unroll
            for (int m = 0; m < R; ++m)
                staged_store_bin<false, LOGM> (sf, j + m * T, logW, v[m]);
            __syncthreads();
            staging_drain<G> (sf, out, j, logW, active);
        }
        else if (active)
        {
#pragma unroll
            for (int m = 0; m < R; ++m)
                reinterpret_cast<float2*> (out)[j + m * T] = v[m];
        }
    }
    else // R2C: split step  X[k] = E - i w_k D,  X[M-k] = conj(E + i w_k D),  E,D = (Z[k] +- Z*[M-k]) / 2
    {
        if (smem_was_read)
            __syncthreads();
        scatter_natural<G> (v, j, s);
        __syncthreads();
#pragma unroll
        for (int m = 0; m < R / 2; ++m)
        {
            const int k = j + m * T;
            v[2 * m] = lds2 (s + pad (k));
            v[2 * m + 1] = lds2 (s + pad ((m == 0 && j == 0) ? M / 2 : M - k));
        }
        if constexpr (UNORD)
            __syncthreads(); // natural-order image fully consumed before the staging image overwrites it
#pragma unroll
        for (int m = 0; m < R / 2; ++m)
        {
            const int k = j + m * T;
            const float2 za = v[2 * m], zb = v[2 * m + 1];
            float2 xa, xb;
            int ka = k, kb = M - k;
            if (m == 0 && j == 0)
            {
                xa = make_float2 (za.x + za.y, za.x - za.y); // (DC, Nyquist)
                xb = make_float2 (zb.x, -zb.y);              // X[M/2] = conj Z[M/2]
                kb = M / 2;
            }
            else
            {
                const float2 w = __ldg (a.rtw + k);
                const float2 e = make_float2 (0.5f * (za.x + zb.x), 0.5f * (za.y - zb.y));
                const float2 d = make_float2 (0.5f * (za.x - zb.x), 0.5f * (za.y + zb.y));
                const float2 wd = cmul_dir<-1> (d, w);
                xa = make_float2 (e.x + wd.y, e.y - wd.x);
                xb = make_float2 (e.x - wd.y, -e.y - wd.x);
            }
            if constexpr (UNORD)
            {
                staged_store_bin<true, LOGM> (sf, ka, logW, xa);
                staged_store_bin<true, LOGM> (sf, kb, logW, xb);
            }
            else if (active)
            {
                reinterpret_cast<float2*> (out)[ka] = xa;
                reinterpret_cast<float2*> (out)[kb] = xb;
            }
        }
        if constexpr (UNORD)
        {
            __syncthreads();
            staging_drain<G> (sf, out, j, logW, active);
        }
    }
}

// threads per CTA / occupancy targets
template <int LOGM, int R>
struct Launch
{
    using G = Geo<LOGM, R>;
    static constexpr int THREADS = G::T >= 256 ? G::T : 256;
    static constexpr int PER_CTA = THREADS / G::T;
    static constexpr int SMEM_BYTES = PER_CTA * G::SMEM_F2 * 8;
    static constexpr int SMEM_BYTES_UNORD = PER_CTA * G::SMEM_F2_UNORD * 8;
    static constexpr int MIN_BLOCKS = THREADS <= 256 ? 4 : (THREADS <= 512 ? 2 : 1);
};

template <int LOGM, int R, int KIND, bool UNORD>
__global__ void __launch_bounds__ (Launch<LOGM, R>::THREADS, Launch<LOGM, R>::MIN_BLOCKS) fft_kernel (const FftArgs a)
{
    fft_body<LOGM, R, KIND, UNORD> (a);
}

// ---------------------------------------------------------------------------------------------
// host side: twiddle tables (fp64 -> fp32), shared by the plan and the emulator tests
// ---------------------------------------------------------------------------------------------
template <int LOGM, int R>
inline void fill_stage_twiddles (float2* tw) // Geo<LOGM,R>::TW_LEN entries
{
    using G = Geo<LOGM, R>;
    for (int s = 1; s < G::S; ++s)
    {
        const int r = G::radix (s), Ns = G::ns (s);
        float2* t = tw + G::tw_off (s);
        for (int q = 1; q < r; ++q)
            for (int k = 0; k < Ns; ++k)
            {
                // exp(-2 pi i k q / (Ns r)) with the angle reduced exactly in integers first
                const long long num = ((long long) k * q) % ((long long) Ns * r);
                const long double ang = -2.0L * 3.141592653589793238462643383279502884L * (long double) num / (long double) ((long long) Ns * r);
                t[(q - 1) * Ns + k] = make_float2 ((float) cosl (ang), (float) sinl (ang));
            }
    }
}
inline void fill_real_twiddles (float2* rtw, int M) // M/2 entries: exp(-2 pi i k / 2M)
{
    for (int k = 0; k < M / 2; ++k)
    {
        const long double ang = -2.0L * 3.141592653589793238462643383279502884L * (long double) k / (long double) (2LL * M);
        rtw[k] = make_float2 ((float) cosl (ang), (float) sinl (ang));
    }
}
} // namespace cfb
