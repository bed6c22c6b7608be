// Large transforms (more points than one CTA can hold): multi-pass "four-step" FFT built from TILE kernels.
//
// The reference runs any size with the same FFTPACK passes over the whole buffer
// (/root/reference/simd/chowdsp_fft_impl_avx.cpp:430-490, cfftf1_ps) using the caller's `work` array as the
// ping-pong buffer (:1861-1863).  Here N = L1 * L2 * L3 (each factor <= 1024) is done in two or three
// passes, each of which reads and writes the array exactly once with full-sector accesses:
//
//   pass A  length-L1 FFTs down the columns of the [L1][N/L1] view (stride N/L1), times W_N^(k1*s), in place
//   pass B  (three-pass plans) for every row k1: length-L2 column FFTs of its [L2][L3] view, times
//           W_(L2 L3)^(k2*n3), in place
//   pass C  length-L3 FFTs of contiguous rows, written transposed:  X[k1 + L1*(k2 + L2*k3)]
//
// A tile kernel owns C (= 8) transforms that are ADJACENT IN MEMORY across the transform index, so the
// strided side of every pass still moves C*8 = 64 contiguous bytes per element row.  The butterflies,
// twiddle scheme and exchange code are the ones of the single-kernel transform (fft_kernels.cuh).
#pragma once
#include "fft_kernels.cuh"
#include "pipe_kernels.cuh" // mbarrier + bulk-copy helpers


namespace cfb
{
// float2 slots between the exchange regions of adjacent transforms of a tile.  Adjacent threads of a warp own
// adjacent transforms, so the region stride must spread the C transforms over all 32 banks: stride mod 16
// float2 slots = an odd multiple of 16 / C (checked with the emulator's bank model for every pass length).
FFT_CX int tile_region_stride (int smem_f2, int C)
{
    int rs = smem_f2;
    while ((rs % 16) % (2 * (16 / C)) != 16 / C)
        ++rs;
    return rs;
}

struct TileArgs
{
    const float2* in;
    float2* out;
    // tile g = blockIdx.x:  base = (g / gdiv) * g_hi + (g % gdiv) * g_lo ;  element idx of transform lt sits at
    // base + lt * tstride + idx * estride   (all in float2 units)
    long long in_g_hi, in_g_lo, in_tstride, in_estride;
    long long out_g_hi, out_g_lo, out_tstride, out_estride;
    int gdiv;
    int ntiles;          // tiles of ONE transform
    int batch;           // transforms in this launch: grid = batch * ntiles
    long long in_bstride, out_bstride; // float2 between consecutive transforms of the batch
    // optional two-level element stride on the input side (distributed exchange layout): element idx sits at
    // (idx >> in_split_log) * in_chunk_stride + (idx & mask) * in_estride ; in_split_log = 31 disables it
    int in_split_log;
    long long in_chunk_stride;
    unsigned tw_c_base;  // added to the twiddle column index (this rank's first global column)
    // four-step twiddle after the transform: out[k] *= W_N^(k * c * tw_mult), c = (g % gdiv) * C + lt,
    // W_N^e = tw_lo[e & mask] * tw_hi[e >> tw_lobits]   (tw_mult == 0: none)
    unsigned tw_mult;
    int tw_lobits;
    const float2* tw_lo;
    const float2* tw_hi;
    const float2* tw; // stage twiddles of the length-L transform
    // Fused exchange of the distributed four-step (phase 0): output row k = element index of the transform goes to
    // rank h = k >> peer_row_log, whose receive buffer is mapped at peer_out[h] (peer memory over NVLink, or this
    // rank's own buffer for h = rank):  peer_out[h] + peer_chunk_off + tile offset + (k & mask) * out_estride.
    // peer_row_log < 0: plain store to `out`.
    int peer_row_log;
    long long peer_chunk_off;
    float2* peer_out[8];
    // L2 prefetch distance in tiles (0 = off): CTA b asks L2 for the input rows of tile b + pf_ahead, about one wave of
    // resident CTAs ahead, so that the strided (DRAM-page-missing) gather of a later CTA finds its lines in L2
    int pf_ahead;
};

template <int DIR>
FFT_HD float2 big_twiddle (const TileArgs& a, unsigned e)
{
    const float2 lo = __ldg (a.tw_lo + (e & ((1u << a.tw_lobits) - 1u)));
    const float2 hi = __ldg (a.tw_hi + (e >> a.tw_lobits));
    return cmul_dir<-1> (lo, hi); // forward twiddle; the caller conjugates for DIR > 0
}

// v[m] = X[jB + m T] of transform ltB -> four-step twiddle -> store (or peer store), shared by tile_body and tile_pipe_body
template <int LOGL, int C, int DIR>
FFT_HD void tile_epilogue (const TileArgs& a, float2 (&v)[16], int ghi, int glo, int ltB, int jB, unsigned cT, const float2* sTw, float2* __restrict__ out)
{
    constexpr int R = 16;
    using G = Geo<LOGL, R>;
    constexpr int T = G::T;
    // v[m] = X[jB + m T] of transform ltB.  Four-step twiddle W_N^(mu k c), k = jB + m T, c = c0 + ltB:
    //   W^(mu k c) = W^(mu jB c) * W^(mu m T c) = B (own register, one table lookup per thread)
    //                                            * A[m][ltB] (16 C values per CTA, looked up once, kept in smem)
    // so a thread makes 2 (+2 for the first 16 C threads) scattered table loads instead of one pair per element.
    if (a.tw_mult != 0)
    {
        const float2 bw = big_twiddle<DIR> (a, (unsigned) jB * cT * a.tw_mult);
#pragma unroll
        for (int m = 0; m < R; ++m)
        {
            const float2 wm = m == 0 ? bw : cmul_dir<-1> (bw, lds2 (sTw + m * C + ltB));
            v[m] = cmul_dir<DIR> (v[m], wm);
        }
    }
    if (a.peer_row_log >= 0)
    {
        // the all-to-all of the distributed transform, done by the stores themselves: every row block goes straight
        // into its owner's receive buffer, so the NVLink transfer overlaps the butterflies of the other tiles
        const long long toff = a.peer_chunk_off + ghi * a.out_g_hi + glo * a.out_g_lo + ltB * a.out_tstride;
        const int mask = (1 << a.peer_row_log) - 1;
#pragma unroll
        for (int m = 0; m < R; ++m)
        {
            const int k = jB + m * T;
            a.peer_out[k >> a.peer_row_log][toff + (long long) (k & mask) * a.out_estride] = v[m];
        }
        return;
    }
    float2* __restrict__ q = out + ltB * a.out_tstride + jB * a.out_estride;
#pragma unroll
    for (int m = 0; m < R; ++m)
        q[(long long) (m * T) * a.out_estride] = v[m];
}

// LOAD_J_FAST: the transforms are contiguous rows (pass C), so the LOAD uses the thread map of the batched
// kernel (consecutive threads = consecutive elements of one row) and the map switches to "consecutive
// threads = adjacent transforms" at the first exchange, which is what makes the transposed store coalesced.
template <int LOGL, int C, int DIR, bool LOAD_J_FAST>
FFT_HD void tile_body (const TileArgs& a)
{
    constexpr int R = 16;
    using G = Geo<LOGL, R>;
    constexpr int T = G::T;
    constexpr int RS = tile_region_stride (G::SMEM_F2, C);
    static_assert (! LOAD_J_FAST || G::S >= 2, "the thread-map switch needs at least one exchange");
    FFT_DYN_SMEM (float2, smem);
    const int tid = (int) threadIdx.x;
    const int bx = (int) blockIdx.x / a.ntiles;
    const int g = (int) blockIdx.x - bx * a.ntiles;
    if (bx >= a.batch)
        return;
    const int ghi = g / a.gdiv, glo = g - ghi * a.gdiv;
    const float2* __restrict__ in = a.in + bx * a.in_bstride + ghi * a.in_g_hi + glo * a.in_g_lo;
    float2* __restrict__ out = a.out + bx * a.out_bstride + ghi * a.out_g_hi + glo * a.out_g_lo;

    const int ltB = tid % C, jB = tid / C; // adjacent threads = adjacent transforms
    const int ltA = tid / T, jA = tid % T; // adjacent threads = adjacent elements
    if (a.pf_ahead > 0)
    {
        const long long tn = (long long) blockIdx.x + a.pf_ahead;
        if (tn < (long long) a.ntiles * a.batch)
        {
            const int bn = (int) (tn / a.ntiles);
            const int gn = (int) (tn - (long long) bn * a.ntiles);
            const int ghn = gn / a.gdiv, gln = gn - ghn * a.gdiv;
            const float2* pn = a.in + bn * a.in_bstride + ghn * a.in_g_hi + gln * a.in_g_lo;
            if constexpr (LOAD_J_FAST)
            {
                // C contiguous rows of L values = L / 16 lines each: line (tid % (L/16)) of row (tid / (L/16))
                constexpr int LPR = G::M / 16;
                for (int i = tid; i < C * LPR; i += T * C)
                    prefetch_l2 (pn + (long long) (i / LPR) * a.in_tstride + (long long) (i % LPR) * 16);
            }
            else
            {
                const int mask = a.in_split_log >= 31 ? -1 : (1 << a.in_split_log) - 1;
                for (int idx = tid; idx < G::M; idx += T * C) // one (C * 8)-byte piece per element row
                    prefetch_l2 (pn + (a.in_split_log >= 31 ? (long long) idx * a.in_estride
                                                              : (long long) (idx >> a.in_split_log) * a.in_chunk_stride + (long long) (idx & mask) * a.in_estride));
            }
        }
    }
    float2 v[R];
    float2* sB = smem + ltB * RS;
    float2* sTw = smem + C * RS;           // A[m][lt] = W_N^(mu m T c(lt)), see the twiddle step below
    const unsigned cT = a.tw_c_base + (unsigned) (glo * C + ltB); // this thread's column in the twiddle index
    if (a.tw_mult != 0)
    {
        for (int i = tid; i - tid < R * C; i += T * C) // i = m C + lt
        {
            if (i < R * C)
            {
                const unsigned c = a.tw_c_base + (unsigned) (glo * C + i % C);
                sts2 (sTw + i, big_twiddle<DIR> (a, (unsigned) (i / C) * (unsigned) T * c * a.tw_mult));
            }
            else
                smem_skip();
        }
    } // visibility: every pass has at least one exchange barrier before the twiddle step

    if constexpr (LOAD_J_FAST)
    {
        const float2* __restrict__ p = in + ltA * a.in_tstride + jA * a.in_estride;
#pragma unroll
        for (int m = 0; m < R; ++m)
            v[m] = ldg_stream (p + (long long) (m * T) * a.in_estride);
        float2* sA = smem + ltA * RS;
        stage_compute<G, DIR, 0> (v, jA, a.tw);
        stage_scatter<G, 0> (v, jA, sA);
        __syncthreads();
        gather_natural<G, 0, R> (v, jB, sB);
        Stages<G, DIR, 1>::run (v, jB, sB, a.tw, true);
    }
    else
    {
        const float2* __restrict__ p = in + ltB * a.in_tstride;
        if (a.in_split_log >= 31)
        {
            p += jB * a.in_estride;
#pragma unroll
            for (int m = 0; m < R; ++m)
                v[m] = ldg_stream (p + (long long) (m * T) * a.in_estride);
        }
        else
        {
            const int mask = (1 << a.in_split_log) - 1;
#pragma unroll
            for (int m = 0; m < R; ++m)
            {
                const int idx = jB + m * T;
                v[m] = ldg_stream (p + (long long) (idx >> a.in_split_log) * a.in_chunk_stride + (long long) (idx & mask) * a.in_estride);
            }
        }
        Stages<G, DIR, 0>::run (v, jB, sB, a.tw, false);
    }

    tile_epilogue<LOGL, C, DIR> (a, v, ghi, glo, ltB, jB, cT, sTw, out);
}

template <int LOGL, int C>
struct TileLaunch
{
    using G = Geo<LOGL, 16>;
    static constexpr int THREADS = G::T * C;
    static constexpr int SMEM_BYTES = (C * tile_region_stride (G::SMEM_F2, C) + 16 * C) * 8; // exchange regions + twiddle rows A[16][C]
    static constexpr int MIN_BLOCKS = THREADS <= 256 ? 4 : (THREADS <= 512 ? 2 : 1);
};

template <int LOGL, int C, int DIR, bool LOAD_J_FAST>
__global__ void __launch_bounds__ (TileLaunch<LOGL, C>::THREADS, TileLaunch<LOGL, C>::MIN_BLOCKS) tile_fft_kernel (const TileArgs a)
{
    tile_body<LOGL, C, DIR, LOAD_J_FAST> (a);
}

// ---------------------------------------------------------------------------------------------
// Persistent TMA-staged tile kernel: the same tile transform as tile_fft_kernel, but a resident CTA loops over tiles
// blockIdx.x, blockIdx.x + gridDim.x, ... and the NEXT tile is brought into a landing buffer by the TMA unit while the
// current one is transformed -- one bulk copy per element row of the tile (C contiguous complex values, 64 or 128
// bytes) on the strided passes, one per transform (a contiguous row of L values) on the contiguous-row pass, all
// completing on one mbarrier.  The strided gather therefore costs no LSU instructions and no register scoreboard
// waits (ncu on tile_fft_kernel: long-scoreboard 5.3 and MIO-throttle 3.7 stalled warps per issue, DRAM 56 %).
// The landing image is [element][transform] (strided passes) or [transform][element] (contiguous rows), so the
// stage-0 registers are read from it with unit-stride 64-bit accesses in the thread map that pass uses anyway.
// Needs 16-byte aligned rows (the launcher checks the base pointer and the strides).
// Shared memory: [landing L C float2][exchange regions + twiddle rows as tile_fft_kernel][mbarrier, 16 bytes].
// ---------------------------------------------------------------------------------------------
template <int LOGL, int C>
struct TilePipeLaunch
{
    using TL = TileLaunch<LOGL, C>;
    using G = Geo<LOGL, 16>;
    static constexpr int THREADS = TL::THREADS;
    // contiguous-row pass: rows of short transforms (T < 16 threads) are pitched L + T so that the transforms a
    // half-warp reads from land in different banks
    static constexpr int ROW_PAD = G::T < 16 ? G::T : 0;
    static constexpr int LAND_BYTES = (G::M + ROW_PAD) * C * 8;
    static constexpr int SMEM_BYTES = LAND_BYTES + TL::SMEM_BYTES + 16;
    static constexpr bool FITS = SMEM_BYTES <= 227 * 1024;
    static constexpr int PER_SM_A = (227 * 1024) / (SMEM_BYTES + 1024);
    static constexpr int PER_SM_B = 2048 / THREADS; // 64 registers per thread
    static constexpr int PER_SM = PER_SM_A < PER_SM_B ? (PER_SM_A < 1 ? 1 : PER_SM_A) : PER_SM_B;
};

template <int LOGL, int C, int DIR, bool LOAD_J_FAST>
FFT_HD void tile_pipe_body (const TileArgs& a)
{
    constexpr int R = 16;
    using G = Geo<LOGL, R>;
    using TP = TilePipeLaunch<LOGL, C>;
    constexpr int T = G::T, L = G::M;
    constexpr int RS = tile_region_stride (G::SMEM_F2, C);
    constexpr int NT = T * C;                  // threads per CTA
    constexpr unsigned TILE_BYTES = (unsigned) (L * C * 8);
    constexpr int LP = L + TP::ROW_PAD;        // landing row pitch of the contiguous-row pass
    static_assert (! LOAD_J_FAST || G::S >= 2, "the thread-map switch needs at least one exchange");
    FFT_DYN_SMEM (char, smem_raw);
    float2* land = reinterpret_cast<float2*> (smem_raw);
    float2* smem = reinterpret_cast<float2*> (smem_raw + TP::LAND_BYTES);
    unsigned long long* bar = reinterpret_cast<unsigned long long*> (smem_raw + TP::LAND_BYTES + TP::TL::SMEM_BYTES);
    const int tid = (int) threadIdx.x;
    const int ltB = tid % C, jB = tid / C; // adjacent threads = adjacent transforms
    const int ltA = tid / T, jA = tid % T; // adjacent threads = adjacent elements
    float2* sB = smem + ltB * RS;
    float2* sTw = smem + C * RS;
    const long long tiles = (long long) a.ntiles * a.batch;
    const long long step = (long long) gridDim.x;

    // start the copies of tile t (every thread issues its share of the rows; thread 0 arms the barrier)
    auto fetch = [&] (long long t)
    {
        const int bx = (int) (t / a.ntiles);
        const int g = (int) (t - (long long) bx * a.ntiles);
        const int ghi = g / a.gdiv, glo = g - ghi * a.gdiv;
        const float2* __restrict__ in = a.in + bx * a.in_bstride + ghi * a.in_g_hi + glo * a.in_g_lo;
        if (tid == 0)
            mbar_expect (bar, TILE_BYTES);
        if constexpr (LOAD_J_FAST)
        {
            if (tid < C) // transform `tid` is a contiguous row of L elements
                bulk_copy (land + tid * LP, in + tid * a.in_tstride, (unsigned) L * 8u, bar);
        }
        else
        {
            const int mask = a.in_split_log >= 31 ? -1 : (1 << a.in_split_log) - 1;
#pragma unroll
            for (int i = 0; i < L / NT; ++i) // element row idx: C adjacent transforms = C contiguous values
            {
                const int idx = tid + i * NT;
                const long long off = a.in_split_log >= 31 ? (long long) idx * a.in_estride
                                                           : (long long) (idx >> a.in_split_log) * a.in_chunk_stride + (long long) (idx & mask) * a.in_estride;
                bulk_copy (land + idx * C, in + off, (unsigned) C * 8u, bar);
            }
        }
    };

    long long t = (long long) blockIdx.x;
    if (tid == 0)
        mbar_init (bar);
    __syncthreads();
    if (t < tiles)
        fetch (t);
    for (unsigned it = 0; t < tiles; t += step, ++it)
    {
        const int bx = (int) (t / a.ntiles);
        const int g = (int) (t - (long long) bx * a.ntiles);
        const int ghi = g / a.gdiv, glo = g - ghi * a.gdiv;
        float2* __restrict__ out = a.out + bx * a.out_bstride + ghi * a.out_g_hi + glo * a.out_g_lo;
        const unsigned cT = a.tw_c_base + (unsigned) (glo * C + ltB);
        float2 v[R];
        mbar_wait (bar, it, TILE_BYTES);
        if constexpr (LOAD_J_FAST)
        {
            const float2* p = land + ltA * LP + jA;
#pragma unroll
            for (int m = 0; m < R; ++m)
                v[m] = lds2 (p + m * T);
        }
        else
        {
            const float2* p = land + tid; // element jB + m T of transform ltB sits at (jB + m T) C + ltB = tid + m T C
#pragma unroll
            for (int m = 0; m < R; ++m)
                v[m] = lds2 (p + m * NT);
        }
        __syncthreads(); // the landing buffer is free again; nobody still reads the previous tile's exchange / twiddle rows
        if (t + step < tiles)
            fetch (t + step);
        if (a.tw_mult != 0)
        {
            for (int i = tid; i - tid < R * C; i += NT) // i = m C + lt
            {
                if (i < R * C)
                {
                    const unsigned c = a.tw_c_base + (unsigned) (glo * C + i % C);
                    sts2 (sTw + i, big_twiddle<DIR> (a, (unsigned) (i / C) * (unsigned) T * c * a.tw_mult));
                }
                else
                    smem_skip();
            }
        } // visibility: every pass has at least one exchange barrier before the twiddle step
        if constexpr (LOAD_J_FAST)
        {
            float2* sA = smem + ltA * RS;
            stage_compute<G, DIR, 0> (v, jA, a.tw);
            stage_scatter<G, 0> (v, jA, sA);
            __syncthreads();
            gather_natural<G, 0, R> (v, jB, sB);
            Stages<G, DIR, 1>::run (v, jB, sB, a.tw, true);
        }
        else
            Stages<G, DIR, 0>::run (v, jB, sB, a.tw, false);
        tile_epilogue<LOGL, C, DIR> (a, v, ghi, glo, ltB, jB, cT, sTw, out);
    }
}

template <int LOGL, int C, int DIR, bool LOAD_J_FAST>
__global__ void __launch_bounds__ (TilePipeLaunch<LOGL, C>::THREADS, TilePipeLaunch<LOGL, C>::PER_SM) tile_pipe_kernel (const TileArgs a)
{
    tile_pipe_body<LOGL, C, DIR, LOAD_J_FAST> (a);
}

// ---------------------------------------------------------------------------------------------
// Real transforms of 2M samples on top of an M-point complex transform (M too large for one CTA):
// the split (forward) / merge (backward) step as its own streaming pass.  One thread per pair (k, M-k).
//   forward : z = M-point FFT of the packed samples (ordered, natural) -> X in the pffft packing or the
//             unordered real layout
//   backward: X -> z' for the M-point inverse FFT
// Same formulas as the fused step in fft_core; w_k = exp(-2 pi i k / 2M) from the two-level tables.
// ---------------------------------------------------------------------------------------------
struct RealPassArgs
{
    const float* in;
    float* out;
    long long in_bstride, out_bstride; // floats between consecutive transforms (blockIdx.y)
    int logM;
    int logW;   // 0: ordered pffft packing, 2 / 3: unordered real layout
    int tw_lobits;
    unsigned tw_mult; // W_2M^k = W_N^(k * tw_mult) in the plan's big tables
    const float2* tw_lo;
    const float2* tw_hi;
};

FFT_HD int real_bin_pos_rt (int bin, int logM, int logW) // float offset of the re part (im is W floats later, or +1 ordered)
{
    const int logQ = logM - logW, Q = 1 << logQ;
    const int r = bin >> logQ, mr = bin & (Q - 1);
    const int m = (r & 1) ? ((Q - mr) & (Q - 1)) : mr;
    const int b = m >> logW, lane = m & ((1 << logW) - 1);
    return ((((b << logW) + r) * 2) << logW) + lane;
}
FFT_HD float2 real_load_bin (const float* p, int bin, int logM, int logW)
{
    if (logW == 0)
        return reinterpret_cast<const float2*> (p)[bin];
    const int pos = real_bin_pos_rt (bin, logM, logW);
    return make_float2 (p[pos], p[pos + (1 << logW)]);
}
FFT_HD void real_store_bin (float* p, int bin, int logM, int logW, float2 v)
{
    if (logW == 0)
    {
        reinterpret_cast<float2*> (p)[bin] = v;
        return;
    }
    const int pos = real_bin_pos_rt (bin, logM, logW);
    p[pos] = v.x;
    p[pos + (1 << logW)] = v.y;
}

template <int DIR>
__global__ void __launch_bounds__ (256) real_pass_kernel (const RealPassArgs a)
{
    const int M = 1 << a.logM;
    const int k = (int) (blockIdx.x * blockDim.x + threadIdx.x); // pair index, k <= M/2 - 1 ; k == 0 also does M/2
    if (k >= M / 2)
        return;
    const float* a_in = a.in + (long long) blockIdx.y * a.in_bstride;
    float* a_out = a.out + (long long) blockIdx.y * a.out_bstride;
    const unsigned e = (unsigned) k * a.tw_mult;
    const float2 lo = __ldg (a.tw_lo + (e & ((1u << a.tw_lobits) - 1u))), hi = __ldg (a.tw_hi + (e >> a.tw_lobits));
    const float2 w = cmul_dir<-1> (lo, hi);
    if (DIR < 0)
    {
        // in: natural-order complex z (float2 array); out: half spectrum
        const float2* z = reinterpret_cast<const float2*> (a_in);
        if (k == 0)
        {
            const float2 z0 = z[0], zh = z[M / 2];
            real_store_bin (a_out, 0, a.logM, a.logW, make_float2 (z0.x + z0.y, z0.x - z0.y));
            real_store_bin (a_out, M / 2, a.logM, a.logW, make_float2 (zh.x, -zh.y));
            return;
        }
        const float2 za = z[k], zm = z[M - k];
        const float2 ee = make_float2 (0.5f * (za.x + zm.x), 0.5f * (za.y - zm.y));
        const float2 dd = make_float2 (0.5f * (za.x - zm.x), 0.5f * (za.y + zm.y));
        const float2 wd = cmul_dir<-1> (dd, w);
        real_store_bin (a_out, k, a.logM, a.logW, make_float2 (ee.x + wd.y, ee.y - wd.x));
        real_store_bin (a_out, M - k, a.logM, a.logW, make_float2 (ee.x - wd.y, -ee.y - wd.x));
    }
    else
    {
        float2* z = reinterpret_cast<float2*> (a_out);
        if (k == 0)
        {
            const float2 x0 = real_load_bin (a_in, 0, a.logM, a.logW), xh = real_load_bin (a_in, M / 2, a.logM, a.logW);
            z[0] = make_float2 (x0.x + x0.y, x0.x - x0.y);
            z[M / 2] = make_float2 (2.f * xh.x, -2.f * xh.y);
            return;
        }
        const float2 xa = real_load_bin (a_in, k, a.logM, a.logW), xm = real_load_bin (a_in, M - k, a.logM, a.logW);
        const float2 ee = make_float2 (xa.x + xm.x, xa.y - xm.y);
        const float2 dd = make_float2 (xa.x - xm.x, xa.y + xm.y);
        const float2 wd = cmul_dir<+1> (dd, w);
        z[k] = make_float2 (ee.x - wd.y, ee.y + wd.x);
        z[M - k] = make_float2 (ee.x + wd.y, wd.x - ee.y);
    }
}

// Complex ordered <-> unordered permutation for large N (the single-kernel sizes fuse it; here it is one
// streaming pass: 64-byte chunk of 8 (or 32-byte chunk of 4) bins per thread group).
//   TO_UNORDERED : in = interleaved natural order, out = unordered layout ; else the inverse.
template <bool TO_UNORDERED>
__global__ void __launch_bounds__ (256) complex_reorder_kernel (const float* in, float* out, long long in_bstride, long long out_bstride, int logN, int logW)
{
    in += (long long) blockIdx.y * in_bstride;
    out += (long long) blockIdx.y * out_bstride;
    const long long bin = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    if (bin >= (1LL << logN))
        return;
    const int logL = logN - logW;
    const int r = (int) (bin >> logL);
    const int rem = (int) (bin & ((1LL << logL) - 1));
    const int b = rem >> logW, lane = rem & ((1 << logW) - 1);
    const long long pos = ((((long long) (b << logW) + r) * 2) << logW) + lane;
    if (TO_UNORDERED)
    {
        const float2 v = reinterpret_cast<const float2*> (in)[bin];
        out[pos] = v.x;
        out[pos + (1 << logW)] = v.y;
    }
    else
        reinterpret_cast<float2*> (out)[bin] = make_float2 (in[pos], in[pos + (1 << logW)]);
}

// host: two-level table for W_N^e, e < N = 2^logN:  lo[e & mask] = W_N^(e & mask), hi[e >> lobits] = W_N^((e >> lobits) << lobits)
inline void fill_big_twiddles (float2* lo, float2* hi, int logN, int lobits)
{
    const long double two_pi = 2.0L * 3.141592653589793238462643383279502884L;
    const long long N = 1LL << logN;
    for (long long e = 0; e < (1LL << lobits); ++e)
    {
        const long double ang = -two_pi * (long double) e / (long double) N;
        lo[e] = make_float2 ((float) cosl (ang), (float) sinl (ang));
    }
    for (long long h = 0; h < (N >> lobits); ++h)
    {
        const long double ang = -two_pi * (long double) (h << lobits) / (long double) N;
        hi[h] = make_float2 ((float) cosl (ang), (float) sinl (ang));
    }
}
} // namespace cfb
