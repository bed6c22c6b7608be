// Persistent, TMA-pipelined single-kernel transform for the largest CTA-resident sizes (2^13 and 2^14 complex
// points per transform, i.e. C2C N = 8192 / 16384 and real N = 16384 / 32768).
//
// Why a second kernel for these sizes: a 2^14-point transform fills half of an SM's register file, so
// fft_kernel (fft_kernels.cuh) runs ONE CTA per SM and its load -> compute -> store phases cannot overlap
// (ncu, profiles/r01_ncu_c2c16384.txt: DRAM 49 %, issue slots 25 % busy, long-scoreboard stalls dominate).  Here
// one CTA per SM stays resident and loops over its share of the batch; the NEXT transform's input is fetched by
// the TMA unit (cp.async.bulk global -> shared, completion on an mbarrier) into a landing buffer while the
// current transform is in registers, so HBM reads, butterflies and HBM writes of neighbouring transforms overlap
// inside one SM.
//
// Shared memory (2^14: 218 KB): landing buffer (8 M bytes) + exchange region (4 M bytes + pad) + the last stage's
// twiddle rows.  The transform is 32 x 32 x r_last.  The stage-0 exchange is an ordinary padded 64-bit exchange that
// borrows the (already consumed) landing buffer; the TMA for the next transform is issued right after it.  The
// stage-1 exchange runs while that copy is in flight, so it only has the half-size region: it moves the data in
// two balanced rounds of 16 registers per thread (round_scatter / round_gather).  [A first version moved real and
// imaginary parts separately through the half-size region for both exchanges: twice the LDS/STS instructions,
// MIO-queue bound, 5.0 TB/s at 2^14 -- profiles/r01_pipe_kernel.txt.]
//
// Replaces the same reference pipeline as fft_kernel (simd/chowdsp_fft_impl_avx.cpp:1848-1935); layouts, signs
// and scaling are identical to it (natural-order complex, pffft-packed real spectra).
#pragma once
#include "fft_kernels.cuh"

namespace cfb
{
#ifdef CFB_PIPE_RTW_TABLE // A/B switch (tools/ only)
constexpr bool kPipeDeriveRtw = false;
#else
constexpr bool kPipeDeriveRtw = true;
#endif

template <int LOGM>
struct PipeGeo
{
    using G = Geo<LOGM, 32>;
    static constexpr int M = G::M, T = G::T, R = 32;
    static constexpr int LAND_BYTES = M * 8;                 // one transform, linear (TMA destination)
    static constexpr int XCH_FLOATS = M + (M >> 5);          // (M/2 + M/64) float2: half-spectrum exchange of the real split step;
                                                             // the stage-1 rounds use M/2 float2 of it, stage 0 spills M/32 into it
    static_assert (G::S == 3 && G::radix (1) == 32, "written for 32 x 32 x r_last");
    // last-stage twiddle rows kept in shared memory: q = 1, 2, 3 and 4, 8, .. (the other powers are one register
    // product each, as in the full-radix stages).  Out of L1 they would not fit next to a > 200 KB carve-out, and
    // every L2 round trip stalls one of only four warps per scheduler.
    static constexpr int RL = G::RLAST;
    static constexpr int TW_ROWS = RL <= 4 ? RL - 1 : 3 + (RL / 4 - 1);
    static constexpr int TW_OFFSET = LAND_BYTES + XCH_FLOATS * 4;
    static constexpr int TW1_OFFSET = TW_OFFSET + TW_ROWS * T * 8; // stage-1 rows w^4 .. w^28, 32 entries each
    static constexpr int BAR_OFFSET = TW1_OFFSET + 7 * 32 * 8;
    static constexpr int SMEM_BYTES = BAR_OFFSET + 16;
    static constexpr int CTAS_PER_SM = (2 * (SMEM_BYTES + 1024) <= 227 * 1024 && 2 * T <= 512) ? 2 : 1;
    static_assert (SMEM_BYTES <= 227 * 1024, "landing + exchange buffers exceed the shared memory of an SM");
    static_assert (G::S >= 2 && RL >= 4, "expects a short last stage of radix >= 4");
    static FFT_CX int tw_row_q (int i) { return i < 3 ? i + 1 : 4 * (i - 2); } // smem row i holds power q
    static_assert (G::T % 32 == 0 && G::R == 32, "the addressing below assumes 32 points per thread");
};

// ---------------------------------------------------------------------------------------------
// mbarrier + bulk-copy helpers (sm_90+ PTX; SASS: SYNCS.*, UBLKCP.S.G)
// ---------------------------------------------------------------------------------------------
#ifdef CHOWDSP_EMU
// the emulator runs every CUDA thread as an OS thread: the "TMA" is a memcpy by the issuing thread followed by a
// release increment of the barrier word, the wait spins on it
FFT_HD void mbar_init (unsigned long long* bar) { __atomic_store_n (bar, 0ull, __ATOMIC_RELEASE); }
FFT_HD void bulk_load (void* dst, const void* src, unsigned bytes, unsigned long long* bar)
{
    std::memcpy (dst, src, bytes);
    __atomic_fetch_add (bar, 1ull, __ATOMIC_RELEASE);
}
FFT_HD void mbar_wait (unsigned long long* bar, unsigned it)
{
    while (__atomic_load_n (bar, __ATOMIC_ACQUIRE) <= (unsigned long long) it)
        std::this_thread::yield();
}
#else
FFT_HD unsigned smem_addr (const void* p) { return (unsigned) __cvta_generic_to_shared (p); }
FFT_HD void mbar_init (unsigned long long* bar)
{
    asm volatile ("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr (bar)) : "memory");
    asm volatile ("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// arm the barrier with the byte count, then start the copy; the barrier phase completes when all bytes landed
FFT_HD void bulk_load (void* dst, const void* src, unsigned bytes, unsigned long long* bar)
{
    asm volatile ("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr (bar)), "r"(bytes) : "memory");
    asm volatile ("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                  ::"r"(smem_addr (dst)), "l"(src), "r"(bytes), "r"(smem_addr (bar)) : "memory");
}
FFT_HD void mbar_wait (unsigned long long* bar, unsigned it)
{
    const unsigned addr = smem_addr (bar), parity = it & 1u;
    unsigned done;
    do
    {
        asm volatile ("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                      : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    } while (done == 0);
}
#endif

// ---------------------------------------------------------------------------------------------
// exchange after stage 1 (Ns = 32, radix 32) in TWO balanced rounds through a buffer of M/2 float2: round h moves
// the outputs q in [16 h, 16 h + 16) of every thread.  Position p = w 1024 + q 32 + l (writer thread 32 w + l) maps to
// slot w 512 + (q - 16 h) 32 + l; reader j needs p = j + m T, which lies in round h iff ((m mod D) div (D/2)) == h
// with D = 1024 / T, at slot (m div D) 512 + (m mod D mod D/2) T + j.  Every thread stores 16 and loads 16 values per
// round (so 32 complex registers stay live), all accesses are unit stride across a warp: conflict free, no padding.
// ---------------------------------------------------------------------------------------------
template <class G, int H>
FFT_HD void round_scatter (const float2 (&v)[32], int j, float2* xs)
{
    float2* sb = xs + (j >> 5) * 512 + (j & 31);
#pragma unroll
    for (int q = 0; q < 16; ++q)
        sts2 (sb + q * 32, v[16 * H + q]);
}
template <class G, int H>
FFT_HD void round_gather (float2 (&v)[32], int j, const float2* xs)
{
    constexpr int D = 1024 / G::T;
    static_assert (D >= 2 && G::T * D == 1024, "expects 256 or 512 threads per transform");
    const float2* rb = xs + j;
#pragma unroll
    for (int m = 0; m < 32; ++m)
        if ((m % D) / (D / 2) == H)
            v[m] = lds2 (rb + (m / D) * 512 + ((m % D) % (D / 2)) * G::T);
}

// stage 1 (Ns = 32, radix 32): its ten twiddle rows depend on the lane only (k = j mod 32), so the persistent loop
// fetches them once per kernel instead of once per transform (as L1 loads they were 20 % of all stall samples,
// profiles/r01_pipe_kernel.txt): w^1, w^2, w^3 stay in registers, w^4, w^8, .., w^28 in shared memory (ten
// register pairs would spill under the 128-register cap of a 512-thread CTA)
struct Stage1Tw
{
    float2 w1, w2, w3;
};
template <class G>
FFT_HD Stage1Tw load_stage1_tw (int j, const float2* __restrict__ tw, float2* tw1s)
{
    constexpr int Ns = 32;
    const float2* __restrict__ t = tw + G::tw_off (1) + (j & (Ns - 1));
    Stage1Tw r;
    r.w1 = __ldg (t);
    r.w2 = __ldg (t + Ns);
    r.w3 = __ldg (t + 2 * Ns);
    if (j < Ns)
    {
#pragma unroll
        for (int a4 = 1; a4 < 8; ++a4)
            sts2 (tw1s + (a4 - 1) * Ns + j, __ldg (t + (2 + a4) * Ns));
    }
    return r;
}
template <int DIR>
FFT_HD void pipe_stage1 (float2 (&v)[32], const Stage1Tw& w, const float2* tw1s_lane)
{
    v[1] = cmul_dir<DIR> (v[1], w.w1);
    v[2] = cmul_dir<DIR> (v[2], w.w2);
    v[3] = cmul_dir<DIR> (v[3], w.w3);
#pragma unroll
    for (int a4 = 1; a4 < 8; ++a4)
    {
        const float2 wa = lds2 (tw1s_lane + (a4 - 1) * 32);
        v[4 * a4] = cmul_dir<DIR> (v[4 * a4], wa);
        v[4 * a4 + 1] = cmul_dir<DIR> (v[4 * a4 + 1], cmul_dir<-1> (w.w1, wa));
        v[4 * a4 + 2] = cmul_dir<DIR> (v[4 * a4 + 2], cmul_dir<-1> (w.w2, wa));
        v[4 * a4 + 3] = cmul_dir<DIR> (v[4 * a4 + 3], cmul_dir<-1> (w.w3, wa));
    }
    RegFft<32, DIR, 1>::run (&v[0]);
}

// last (short) stage with the twiddle rows in shared memory: butterfly u of this thread has k = j + u T and
// W^(k q) = W^(j q) W_R^(u q); W^(j q) for q not in the table is one product of two rows
template <class P, int DIR>
FFT_HD void pipe_last_stage (float2 (&v)[32], int j, const float2* tws)
{
    using G = typename P::G;
    constexpr int r = P::RL, SUB = 32 / r, T = G::T;
    float2 w[r];
    const float2* tj = tws + j;
#pragma unroll
    for (int i = 0; i < P::TW_ROWS; ++i)
        w[P::tw_row_q (i)] = lds2 (tj + i * T);
#pragma unroll
    for (int q = 5; q < r; ++q)
        if ((q & 3) != 0)
            w[q] = cmul_dir<-1> (w[q & 3], w[q & ~3]);
#pragma unroll
    for (int q = 1; q < r; ++q)
    {
#pragma unroll
        for (int u = 0; u < SUB; ++u)
        {
            const float2 x = cmul_dir<DIR> (v[u + q * SUB], w[q]);
            v[u + q * SUB] = mul_w32_rt<DIR> (x, u * q);
        }
    }
#pragma unroll
    for (int u = 0; u < SUB; ++u)
        RegFft<r, DIR, SUB>::run (&v[u]);
}

// ---------------------------------------------------------------------------------------------
// the kernel: grid = min (batch, SMs * CTAS_PER_SM), blockDim = T = M / 32, dynamic smem = PipeGeo::SMEM_BYTES.
// Plain batches only (transform x reads in + x in_inner, writes out + x out_inner); the input rows must be
// 16-byte aligned (TMA), which the launcher checks.  Ordered layouts (LOGW = 0).
// ---------------------------------------------------------------------------------------------
template <int LOGM, int KIND>
FFT_HD void pipe_body (const FftArgs& a)
{
    using P = PipeGeo<LOGM>;
    using G = typename P::G;
    constexpr int R = 32, T = G::T, M = G::M;
    constexpr int DIR = (KIND == C2C_FWD || KIND == R2C) ? -1 : +1;
    FFT_DYN_SMEM (char, smem);
    float2* land = reinterpret_cast<float2*> (smem);
    float2* xs = reinterpret_cast<float2*> (smem + P::LAND_BYTES);
    float2* tws = reinterpret_cast<float2*> (smem + P::TW_OFFSET);
    float2* tw1s = reinterpret_cast<float2*> (smem + P::TW1_OFFSET);
    unsigned long long* bar = reinterpret_cast<unsigned long long*> (smem + P::BAR_OFFSET);

    const int j = (int) threadIdx.x;
    {   // last-stage twiddle rows: global table row q-1 (T entries each, see Geo) -> shared row i
        const float2* __restrict__ tl = a.tw + G::tw_off (G::S - 1) + j;
#pragma unroll
        for (int i = 0; i < P::TW_ROWS; ++i)
            sts2 (tws + i * T + j, __ldg (tl + (P::tw_row_q (i) - 1) * T));
    }
    // real split / merge twiddle of this thread's first bin, w_j / 2; the others are w_(j + m T) = w_j W_64^m
    float2 wj = make_float2 (0.f, 0.f);
    if constexpr (KIND == R2C || KIND == C2R)
        wj = __ldg (a.rtw + j);
    const Stage1Tw w1t = load_stage1_tw<G> (j, a.tw, tw1s);
    const long long step = (long long) gridDim.x;
    long long x = (long long) blockIdx.x;
    if (j == 0)
        mbar_init (bar);
    __syncthreads();
    if (j == 0 && x < a.batch)
        bulk_load (land, a.in + x * a.in_inner, (unsigned) P::LAND_BYTES, bar);

    for (unsigned it = 0; x < a.batch; x += step, ++it)
    {
        float2 v[R];
        mbar_wait (bar, it);
        // ---- prologue: stage-0 registers v[m] = input element j + m T, from the landing buffer ----
        if constexpr (KIND != C2R)
        {
            const float2* lj = land + j;
#pragma unroll
            for (int m = 0; m < R; ++m)
                v[m] = lds2 (lj + m * T);
        }
        else
        {
            // merge step  Z'[k] = (X[k] + X*[M-k]) + i conj(w_k) (X[k] - X*[M-k]),  w_k = e^{-2 pi i k / 2M}: both
            // operands come straight from the landing buffer, so each thread forms all of its own k = j + m T.
            // For k >= M/2, w_k = -i w_{k - M/2}: the table only holds k < M/2 (times 1/2).
            const float2* lj = land + j;
            const float2 two = make_float2 (2.f, 2.f);
#pragma unroll
            for (int m = 0; m < R / 2; ++m)
            {
                const float2 wh = real_tw<32, kPipeDeriveRtw> (wj, a.rtw + j + m * T, m);
                {
                    const float2 xa = lds2 (lj + m * T);
                    const float2 xm = lds2 ((m == 0 && j == 0) ? land : land + (M - m * T) - j);
                    const float2 cm = make_float2 (xm.x, -xm.y);
                    const float2 e = f2_add (xa, cm), d = f2_sub (xa, cm);
                    const float2 wd = cmul_dir<+1> (d, wh);                       // conj(w_k) d / 2
                    float2 zk = f2_fma (make_float2 (-wd.y, wd.x), two, e);      // e + i conj(w_k) d
                    if (m == 0 && j == 0)
                        zk = make_float2 (xa.x + xa.y, xa.x - xa.y);             // Z'[0] from (DC, Nyquist)
                    v[m] = zk;
                }
                {
                    const float2 xa = lds2 (lj + (m + R / 2) * T);
                    const float2 xm = lds2 (land + (M / 2 - m * T) - j);          // M - k, k = j + m T + M/2
                    const float2 cm = make_float2 (xm.x, -xm.y);
                    const float2 e = f2_add (xa, cm), d = f2_sub (xa, cm);
                    const float2 wd = cmul_dir<+1> (d, wh);                       // conj(w_{k - M/2}) d / 2
                    v[m + R / 2] = f2_fma (wd, make_float2 (-2.f, -2.f), e);     // i * (i conj(w') d) = - conj(w') d
                }
            }
        }
        // ---- stage 0, full 64-bit exchange through [landing | exchange] (padded, M + M/32 slots) ----
        stage_compute<G, DIR, 0> (v, j, a.tw);
        __syncthreads(); // the landing buffer has been consumed, and the previous iteration is done with xs
        stage_scatter<G, 0> (v, j, land);
        __syncthreads();
        gather_natural<G, 0, R> (v, j, land);
        __syncthreads(); // landing buffer free again: fetch the next transform while this one finishes
        if (j == 0 && x + step < a.batch)
            bulk_load (land, a.in + (x + step) * a.in_inner, (unsigned) P::LAND_BYTES, bar);
        // ---- stage 1, two-round exchange through the exchange region alone ----
        pipe_stage1<DIR> (v, w1t, tw1s + (j & 31));
        {
            float2 n[R]; // the gathered values land in fresh registers: v[16..31] are still needed by round 1
            round_scatter<G, 0> (v, j, xs);
            __syncthreads();
            round_gather<G, 0> (n, j, xs);
            __syncthreads();
            round_scatter<G, 1> (v, j, xs);
            __syncthreads();
            round_gather<G, 1> (n, j, xs);
#pragma unroll
            for (int m = 0; m < R; ++m)
                v[m] = n[m];
        }
        // ---- last stage ----
        pipe_last_stage<P, DIR> (v, j, tws);

        // ---- epilogue ----
        float* out = a.out + x * a.out_inner;
        if constexpr (KIND != R2C)
        {
            float2* __restrict__ out2 = reinterpret_cast<float2*> (out) + j;
#pragma unroll
            for (int m = 0; m < R; ++m)
                out2[m * T] = v[m];
        }
        else
        {
            // split step  X[k] = E - i w_k D,  X[M-k] = conj(E + i w_k D),  E,D = (Z[k] +- Z*[M-k]) / 2.
            // Z[k], k = j + m T < M/2, is register m; Z[M-k] is register R-1-m of thread T-j: the upper half of the
            // spectrum crosses the exchange region (seen as float2, shifted so that bin M/2 sits at slot 0).
            float2* xh = xs - G::pad (M / 2);
            __syncthreads();
            scatter_natural<G, R / 2, R> (v, j, xh);
            __syncthreads();
            float2 zb[R / 2];
#pragma unroll
            for (int m = 0; m < R / 2; ++m)
                zb[m] = lds2 (xh + ((m == 0 && j == 0) ? G::pad (M / 2) : mirror_slot<G> (j, m)));
            float2* __restrict__ lo = reinterpret_cast<float2*> (out) + j;
            float2* __restrict__ hi = reinterpret_cast<float2*> (out) + (M - T) - j;
#pragma unroll
            for (int m = 0; m < R / 2; ++m)
            {
                const float2 za = v[m], zm = zb[m];
                const float2 wh = real_tw<32, kPipeDeriveRtw> (wj, a.rtw + j + m * T, m);
                const float2 cm = make_float2 (zm.x, -zm.y);
                const float2 e = f2_add (za, cm), d = f2_sub (za, cm);
                const float2 wd = cmul_dir<-1> (d, wh);
                float2 xa = f2_fma (e, make_float2 (0.5f, 0.5f), make_float2 (wd.y, -wd.x));
                float2 xm = f2_fma (e, make_float2 (0.5f, -0.5f), make_float2 (-wd.y, -wd.x));
                const bool special = (m == 0 && j == 0);
                if (special)
                {
                    xa = make_float2 (za.x + za.y, za.x - za.y);
                    xm = make_float2 (zm.x, -zm.y);
                }
                lo[m * T] = xa;
                float2* ph = special ? reinterpret_cast<float2*> (out) + M / 2 : hi - m * T + T;
                *ph = xm;
            }
        }
    }
}

template <int LOGM, int KIND>
__global__ void __launch_bounds__ (PipeGeo<LOGM>::T, PipeGeo<LOGM>::CTAS_PER_SM) pipe_kernel (const FftArgs a)
{
    pipe_body<LOGM, KIND> (a);
}
} // namespace cfb
