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
#include "tma.cuh"


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

// L2 cache policy of one side of a pass (TileArgs::in_policy / out_policy).  The chunked plans (large_plan.h) keep the
// intermediate of two consecutive passes in a small ring that is meant to stay L2-resident: ring accesses ask for
// evict_last, the HBM-facing streams of the same kernels for evict_first, so that streaming data does not push the ring out.
enum TilePolicy : int
{
    POLICY_NORMAL = 0,
    POLICY_STREAM = 1, // evict_first
    POLICY_KEEP = 2    // evict_last
};
FFT_HD unsigned long long make_l2_policy (int which)
{
#ifdef CHOWDSP_EMU
    return (unsigned long long) which;
#else
    unsigned long long p;
    if (which == POLICY_STREAM)
        asm volatile ("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    else if (which == POLICY_KEEP)
        asm volatile ("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    else
        asm volatile ("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
    return p;
#endif
}
FFT_HD float2 ldg_hint (const float2* p, unsigned long long policy)
{
#ifdef CHOWDSP_EMU
    (void) policy;
    return *p;
#else
    float2 r;
    asm volatile ("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.f32 {%0, %1}, [%2], %3;" : "=f"(r.x), "=f"(r.y) : "l"(p), "l"(policy));
    return r;
#endif
}
FFT_HD void stg_hint (float2* p, float2 v, unsigned long long policy)
{
#ifdef CHOWDSP_EMU
    (void) policy;
    *p = v;
#else
    asm volatile ("st.global.L2::cache_hint.v2.f32 [%0], {%1, %2}, %3;" ::"l"(p), "f"(v.x), "f"(v.y), "l"(policy) : "memory");
#endif
}

struct TileArgs
{
    // in / out = base of transform 0's buffer (or of a ring slot); this launch's tile 0 starts in_bin0 / out_bin0
    // complex elements into it (chunked launches cover a sub-range of a pass's tiles)
    const float2* in;
    float2* out;
    long long in_bin0, out_bin0;
    // tile g = blockIdx.x:  base = bin0 + (g / gdiv) * g_hi + (g % gdiv) * g_lo ;  element idx of transform lt sits at
    // base + lt * tstride + idx * estride   (all in float2 units)
    long long in_g_hi, in_g_lo, in_tstride, in_estride;
    long long out_g_hi, out_g_lo, out_tstride, out_estride;
    int gdiv;
    int ntiles;          // tiles of ONE transform in this launch
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
    // Unordered complex layouts (the reference's pffft_zreorder permutation, simd/chowdsp_fft_impl_avx.cpp:1816-1838) folded
    // into the addressing of the FIRST pass's loads (UIO = 1, inverse transforms) / the LAST pass's stores (UIO = 2,
    // forward transforms): the element offset computed above is the natural BIN index of a 2^logN-point transform and
    // the access goes to that bin's slot of the W = 2^unord_logW lane layout.  No separate reorder sweep.
    int logN, unord_logW;
    int in_policy, out_policy; // TilePolicy
};

template <int DIR>
FFT_HD float2 big_twiddle (const TileArgs& a, unsigned e)
{
    const float2 lo = __ldg (a.tw_lo + (e & ((1u << a.tw_lobits) - 1u)));
    const float2 hi = __ldg (a.tw_hi + (e >> a.tw_lobits));
    return cmul_dir<-1> (lo, hi); // forward twiddle; the caller conjugates for DIR > 0
}

// float offset, inside an unordered complex spectrum of 2^logN bins, of the 8-byte pair a thread accesses for bin `bin`:
// the even bin of a pair owns (re, re') at the pair's slot, the odd one (im, im') W floats later (SURVEY.md §8a-L:
// vector 2k holds W real parts, vector 2k+1 the W imaginary parts of bins r (N/W) + b W + lane, k = b W + r)
FFT_HD long long unord_pair_offset (long long bin, int logN, int logW)
{
    const int logL = logN - logW;
    const long long r = bin >> logL, rem = bin & ((1LL << logL) - 1);
    const long long blk = rem >> logW, lane = rem & ((1 << logW) - 2);
    return ((((blk << logW) + r) * 2) << logW) + lane + ((bin & 1) << logW);
}

// v[m] = X[jB + m T] of transform ltB -> four-step twiddle -> store (or peer store)
// obase = the transform's output buffer, tbin = bin0 + tile offset (element offset of the tile's first transform)
// (natural-order outputs pass obase = buffer + tile offset and tbin = 0: one live 64-bit value instead of two)
template <int LOGL, int C, int DIR, int UIO, int R>
FFT_HD void tile_epilogue (const TileArgs& a, float2 (&v)[R], long long tbin, int ltB, int jB, unsigned cT, const float2* sTw, float2* __restrict__ obase)
{
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
        const long long toff = a.peer_chunk_off + tbin + ltB * a.out_tstride;
        const int mask = (1 << a.peer_row_log) - 1;
#pragma unroll
        for (int m = 0; m < R; ++m)
        {
            const int k = jB + m * T;
            a.peer_out[k >> a.peer_row_log][toff + (long long) (k & mask) * a.out_estride] = v[m];
        }
        return;
    }
    const unsigned long long pol = make_l2_policy (a.out_policy);
    if constexpr (UIO == 2)
    {
        // unordered output: lanes (bins) 2i and 2i+1 are adjacent threads; they swap one float so that the even one holds
        // (re, re') and the odd one (im, im') -- contiguous 8-byte pairs of the unordered layout (as fft_core's UDIRECT)
        // The element index k = jB + m T is the MOST significant digit of the bin (bin = tile offset + k * estride, estride =
        // 2^(logN - LOGL)), and m is its top log2(R) bits: the lane row r of the unordered layout is the top logW bits of m, the
        // rest of m stays in the in-row index -- so every store address is (one per-thread base) + (two small multiples of m).
        float* __restrict__ fb = reinterpret_cast<float*> (obase);
        const long long bin0 = tbin + ltB * a.out_tstride + jB * a.out_estride;
        const int odd = ltB & 1;
        constexpr int LOGR = ilog2 (R);
        const int sh = LOGR - a.unord_logW;
        const long long step_lo = 1LL << (a.logN - LOGR + a.unord_logW + 1); // floats per unit of (m mod 2^sh)
        const int step_r = 2 << a.unord_logW;                                 // floats per lane row
        float* __restrict__ pb = fb + unord_pair_offset (bin0, a.logN, a.unord_logW);
#pragma unroll
        for (int m = 0; m < R; ++m)
        {
            const float recv = shfl1 (odd ? v[m].x : v[m].y, ((int) threadIdx.x ^ 1) & 31, 32);
            const float2 o = odd ? make_float2 (recv, v[m].y) : make_float2 (v[m].x, recv);
            stg_hint (reinterpret_cast<float2*> (pb + (long long) (m & ((1 << sh) - 1)) * step_lo + (m >> sh) * step_r), o, pol);
        }
        return;
    }
    float2* __restrict__ q = obase + tbin + ltB * a.out_tstride + jB * a.out_estride;
#pragma unroll
    for (int m = 0; m < R; ++m)
        stg_hint (q + (long long) (m * T) * a.out_estride, v[m], pol);
}

// LOAD_J_FAST: the transforms are contiguous rows (pass C), so the LOAD uses the thread map of the batched
// kernel (consecutive threads = consecutive elements of one row) and the map switches to "consecutive
// threads = adjacent transforms" at the first exchange, which is what makes the transposed store coalesced.
// UIO: 0 = natural-order interleaved complex on both sides, 1 = unordered input (first pass of an inverse transform;
// strided passes only), 2 = unordered output (last pass of a forward transform)
// R: complex points per thread -- 16 (three Stockham stages at 512 / 1024 points, 64 registers, 512-thread CTAs) or 32 (two
// stages, one shared-memory exchange fewer, 128 registers, 256-thread CTAs; tuning hook tile_r)
template <int LOGL, int C, int DIR, bool LOAD_J_FAST, int UIO = 0, int R = 16>
FFT_HD void tile_body (const TileArgs& a)
{
    using G = Geo<LOGL, R>;
    constexpr int T = G::T;
    constexpr int RS = tile_region_stride (G::SMEM_F2, C);
    static_assert (! LOAD_J_FAST || G::S >= 2, "the thread-map switch needs at least one exchange");
    static_assert (UIO != 1 || ! LOAD_J_FAST, "unordered inputs enter through a strided (first) pass");
    FFT_DYN_SMEM (float2, smem);
    const int tid = (int) threadIdx.x;
    const int bx = (int) blockIdx.x / a.ntiles;
    const int g = (int) blockIdx.x - bx * a.ntiles;
    if (bx >= a.batch)
        return;
    const int ghi = g / a.gdiv, glo = g - ghi * a.gdiv;
    const float2* __restrict__ ibase = a.in + bx * a.in_bstride;
    const long long in_tbin = a.in_bin0 + ghi * a.in_g_hi + glo * a.in_g_lo;
    const long long out_tbin_full = a.out_bin0 + ghi * a.out_g_hi + glo * a.out_g_lo;
    const float2* __restrict__ in = ibase + in_tbin;
    // unordered outputs address by bin; the others by pointer
    const long long out_tbin = (UIO == 2 || a.peer_row_log >= 0) ? out_tbin_full : 0;
    float2* __restrict__ obase = a.out + bx * a.out_bstride + ((UIO == 2 || a.peer_row_log >= 0) ? 0 : out_tbin_full);

    const int ltB = tid % C, jB = tid / C; // adjacent threads = adjacent transforms
    const int ltA = tid / T, jA = tid % T; // adjacent threads = adjacent elements
    const unsigned long long ipol = make_l2_policy (a.in_policy);
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
            v[m] = ldg_hint (p + (long long) (m * T) * a.in_estride, ipol);
        float2* sA = smem + ltA * RS;
        stage_compute<G, DIR, 0> (v, jA, a.tw);
        stage_scatter<G, 0> (v, jA, sA);
        __syncthreads();
        gather_natural<G, 0, R> (v, jB, sB);
        Stages<G, DIR, 1>::run (v, jB, sB, a.tw, true);
    }
    else
    {
        if constexpr (UIO == 1)
        {
            // unordered input: the even bin of a pair loads (re, re'), the odd one (im, im'); one shuffle each way
            const float* __restrict__ fb = reinterpret_cast<const float*> (ibase);
            const long long bin0 = in_tbin + ltB * a.in_tstride + jB * a.in_estride;
            const int odd = ltB & 1;
            // as in the unordered stores of tile_epilogue: the element index is the top digit of the bin, m its top bits
            constexpr int LOGR = ilog2 (R);
            const int sh = LOGR - a.unord_logW;
            const long long step_lo = 1LL << (a.logN - LOGR + a.unord_logW + 1);
            const int step_r = 2 << a.unord_logW;
            const float* __restrict__ pb = fb + unord_pair_offset (bin0, a.logN, a.unord_logW);
#pragma unroll
            for (int m = 0; m < R; ++m)
            {
                const float2 ld = ldg_hint (reinterpret_cast<const float2*> (pb + (long long) (m & ((1 << sh) - 1)) * step_lo + (m >> sh) * step_r), ipol);
                const float recv = shfl1 (odd ? ld.x : ld.y, (tid ^ 1) & 31, 32);
                v[m] = odd ? make_float2 (recv, ld.y) : make_float2 (ld.x, recv);
            }
        }
        else
        {
            const float2* __restrict__ p = in + ltB * a.in_tstride;
            if (a.in_split_log >= 31)
            {
                p += jB * a.in_estride;
#pragma unroll
                for (int m = 0; m < R; ++m)
                    v[m] = ldg_hint (p + (long long) (m * T) * a.in_estride, ipol);
            }
            else
            {
                const int mask = (1 << a.in_split_log) - 1;
#pragma unroll
                for (int m = 0; m < R; ++m)
                {
                    const int idx = jB + m * T;
                    v[m] = ldg_hint (p + (long long) (idx >> a.in_split_log) * a.in_chunk_stride + (long long) (idx & mask) * a.in_estride, ipol);
                }
            }
        }
        Stages<G, DIR, 0>::run (v, jB, sB, a.tw, false);
    }

    tile_epilogue<LOGL, C, DIR, UIO, R> (a, v, out_tbin, ltB, jB, cT, sTw, obase);
}

template <int LOGL, int C, int R = 16>
struct TileLaunch
{
    using G = Geo<LOGL, R>;
    static constexpr int THREADS = G::T * C;
    static constexpr int SMEM_BYTES = (C * tile_region_stride (G::SMEM_F2, C) + R * C) * 8; // exchange regions + twiddle rows A[R][C]
    static constexpr int REG_THREADS = R == 16 ? 1024 : 512;                                 // resident threads per SM at 64 / 128 registers
    static constexpr int MIN_BLOCKS = REG_THREADS / THREADS < 1 ? 1 : (REG_THREADS / THREADS > 4 ? 4 : REG_THREADS / THREADS);
};

template <int LOGL, int C, int DIR, bool LOAD_J_FAST, int UIO, int R = 16>
__global__ void __launch_bounds__ (TileLaunch<LOGL, C, R>::THREADS, TileLaunch<LOGL, C, R>::MIN_BLOCKS) tile_fft_kernel (const TileArgs a)
{
    tile_body<LOGL, C, DIR, LOAD_J_FAST, UIO, R> (a);
}

// ---------------------------------------------------------------------------------------------
// Persistent tile kernel with tensor-map TMA on both sides (round 2).  ncu on tile_fft_kernel (profiles/r02_tile_passes.txt):
// DRAM 56..59 %, issue slots 40 %, L1 60 % -- nothing saturated; the top stall is the long scoreboard of the strided gather
// (4.5..4.9 warps per issue): two 512-thread CTAs per SM that load -> compute -> store in near lock step do not keep enough
// bytes in flight.  Here ONE resident CTA per SM loops over tiles; a tile arrives in a landing slot by ONE
// cp.async.bulk.tensor copy per 256 rows (a box of C contiguous values x rows -- the per-row bulk copies of round 1's
// tile_pipe_kernel were TMA-issue bound), the NEXT tile's copy is issued before the current tile is touched (two slots),
// and the result leaves the same way: staged as the tile's output image in the slot it came from, written by one
// tensor-map store per 256 rows.  No LSU instruction touches global memory; 128 KB of loads stay in flight per SM.
// Same butterflies, exchange layout and four-step twiddle as tile_fft_kernel; natural-order data on both sides, no peer
// stores, no split input rows (those cases keep tile_fft_kernel).
// Shared memory: [slot 0][slot 1][exchange regions + twiddle rows as tile_fft_kernel][2 mbarriers].
// ---------------------------------------------------------------------------------------------
struct TileTmaCoords
{
    // coordinates of tile (ghi, glo) of batch element bx, per tensor-map dimension: c = base + ghi * per_hi + glo * per_lo
    // (+ bx in the batch dimension, + the row / chunk index the kernel iterates over in `iter_dim`)
    int per_hi[5], per_lo[5];
    int batch_dim, iter_dim;
    int iters, iter_step; // copies per tile and the coordinate step between them (rows per box)
};

template <int LOGL, int C, int R = 16>
struct TileTmaLaunch
{
    using TL = TileLaunch<LOGL, C, R>;
    using G = Geo<LOGL, R>;
    static constexpr int THREADS = TL::THREADS;
    static constexpr int SLOT_BYTES = G::M * C * 8;
    static constexpr int XCH_OFFSET = 2 * SLOT_BYTES;
    static constexpr int BAR_OFFSET = XCH_OFFSET + TL::SMEM_BYTES;
    static constexpr int SMEM_BYTES = BAR_OFFSET + 16;
    static constexpr bool FITS = SMEM_BYTES <= 227 * 1024 && THREADS <= 1024;
    static constexpr int ROWS_PER_BOX = G::M < 256 ? G::M : 256; // strided side: element rows per tensor copy
};

template <int LOGL, int C, int DIR, bool LOAD_J_FAST, int R>
FFT_HD void tile_tma_body (const TensorMap5* imap, const TensorMap5* omap, const TileArgs& a, const TileTmaCoords& ic, const TileTmaCoords& oc)
{
    using G = Geo<LOGL, R>;
    using TT = TileTmaLaunch<LOGL, C, R>;
    constexpr int T = G::T, L = G::M;
    constexpr int RS = tile_region_stride (G::SMEM_F2, C);
    constexpr int NT = T * C;
    constexpr unsigned TILE_BYTES = (unsigned) TT::SLOT_BYTES;
    static_assert (! LOAD_J_FAST || G::S >= 2, "the thread-map switch needs at least one exchange");
    FFT_DYN_SMEM (char, smem_raw);
    float2* slot[2] = { reinterpret_cast<float2*> (smem_raw), reinterpret_cast<float2*> (smem_raw + TT::SLOT_BYTES) };
    float2* smem = reinterpret_cast<float2*> (smem_raw + TT::XCH_OFFSET);
    unsigned long long* bar = reinterpret_cast<unsigned long long*> (smem_raw + TT::BAR_OFFSET);
    const int tid = (int) threadIdx.x;
    const int ltB = tid % C, jB = tid / C; // adjacent threads = adjacent transforms
    const int ltA = tid / T, jA = tid % T; // adjacent threads = adjacent elements
    float2* sB = smem + ltB * RS;
    float2* sTw = smem + C * RS;
    const long long tiles = (long long) a.ntiles * a.batch;
    const long long step = (long long) gridDim.x;

    auto coords = [&] (const TileTmaCoords& k, long long t, int it, int (&c)[5])
    {
        const int bx = (int) (t / a.ntiles);
        const int g = (int) (t - (long long) bx * a.ntiles);
        const int ghi = g / a.gdiv, glo = g - ghi * a.gdiv;
#pragma unroll
        for (int d = 0; d < 5; ++d)
            c[d] = ghi * k.per_hi[d] + glo * k.per_lo[d];
        c[k.batch_dim] += bx;
        c[k.iter_dim] += it * k.iter_step;
    };
    auto fetch = [&] (long long t, int s)
    {
        mbar_expect (bar + s, TILE_BYTES);
        for (int it = 0; it < ic.iters; ++it)
        {
            int c[5];
            coords (ic, t, it, c);
            tma_load_5d (reinterpret_cast<char*> (slot[s]) + (size_t) it * (TILE_BYTES / (unsigned) ic.iters), imap, c[0], c[1], c[2], c[3], c[4], bar + s);
        }
    };

    long long t = (long long) blockIdx.x;
    if (tid == 0)
    {
        mbar_init (bar);
        mbar_init (bar + 1);
    }
    __syncthreads();
    if (tid == 0 && t < tiles)
        fetch (t, 0);
    for (unsigned it = 0; t < tiles; t += step, ++it)
    {
        const int s = (int) (it & 1u);
        if (tid == 0)
        {
            tma_store_wait_read(); // the previous tile's store has read slot s ^ 1: it may receive the next tile
            if (t + step < tiles)
                fetch (t + step, s ^ 1);
        }
        const int bx = (int) (t / a.ntiles);
        const int g = (int) (t - (long long) bx * a.ntiles);
        const int glo = g % a.gdiv;
        const unsigned cT = a.tw_c_base + (unsigned) (glo * C + ltB);
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
        float2 v[R];
        mbar_wait (bar + s, it >> 1, TILE_BYTES);
        if constexpr (LOAD_J_FAST)
        {
            const float2* p = slot[s] + ltA * L + jA; // landing image [transform][element]
#pragma unroll
            for (int m = 0; m < R; ++m)
                v[m] = lds2 (p + m * T);
            float2* sA = smem + ltA * RS;
            stage_compute<G, DIR, 0> (v, jA, a.tw);
            stage_scatter<G, 0> (v, jA, sA);
            __syncthreads();
            gather_natural<G, 0, R> (v, jB, sB);
            Stages<G, DIR, 1>::run (v, jB, sB, a.tw, true);
        }
        else
        {
            const float2* p = slot[s] + tid; // landing image [element][transform]: (jB + m T) C + ltB = tid + m T C
#pragma unroll
            for (int m = 0; m < R; ++m)
                v[m] = lds2 (p + m * NT);
            Stages<G, DIR, 0>::run (v, jB, sB, a.tw, false);
        }
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
        // output image [element k = jB + m T][transform ltB] in the slot the tile came from (every thread's reads of the
        // slot are behind at least one CTA barrier of the stages above)
        {
            float2* q = slot[s] + tid;
#pragma unroll
            for (int m = 0; m < R; ++m)
                sts2 (q + m * NT, v[m]);
        }
        fence_proxy_async();
        __syncthreads();
        if (tid == 0)
        {
            for (int io = 0; io < oc.iters; ++io)
            {
                int c[5];
                coords (oc, t, io, c);
                tma_store_5d (reinterpret_cast<const char*> (slot[s]) + (size_t) io * (TILE_BYTES / (unsigned) oc.iters), omap, c[0], c[1], c[2], c[3], c[4]);
            }
            tma_store_commit();
        }
    }
    if (tid == 0)
        tma_store_wait_all();
}

#ifndef CHOWDSP_EMU
template <int LOGL, int C, int DIR, bool LOAD_J_FAST, int R = 16>
__global__ void __launch_bounds__ (TileTmaLaunch<LOGL, C, R>::THREADS, 1) tile_tma_kernel (CFB_TMAP5_PARAM imap, CFB_TMAP5_PARAM omap, const TileArgs a, const TileTmaCoords ic, const TileTmaCoords oc)
{
    tile_tma_body<LOGL, C, DIR, LOAD_J_FAST, R> (&imap, &omap, a, ic, oc);
}
#else
template <int LOGL, int C, int DIR, bool LOAD_J_FAST, int R = 16>
void tile_tma_kernel (CFB_TMAP5_PARAM imap, CFB_TMAP5_PARAM omap, const TileArgs a, const TileTmaCoords ic, const TileTmaCoords oc)
{
    tile_tma_body<LOGL, C, DIR, LOAD_J_FAST, R> (&imap, &omap, a, ic, oc);
}
#endif

// ---------------------------------------------------------------------------------------------
// tile_fft_kernel + tensor-map L2 prefetch (tuning hook "tile_pf" = distance in tiles).  ncu on tile_fft_kernel: two resident
// CTAs per SM keep at most 128 KB of loads in flight per SM and only while they wait (about a quarter of a CTA's life), and the
// strided gather's latency is 3-4 us -- by Little's law that is the 2.3 TB/s of reads the passes reach, while a copy kernel with
// the same access pattern and twice the resident threads reaches 2.9.  Here thread 0 of CTA b asks L2 for the input tile of CTA
// b + distance with ONE cp.async.bulk.prefetch.tensor per 256 element rows (no registers, no shared memory, no LSU slot), so
// that the tile is an L2 hit when its CTA starts.  Natural-order passes whose input side a tensor map can describe
// (large_plan.h: build_tile_tma); the others run tile_fft_kernel.
// ---------------------------------------------------------------------------------------------
template <int LOGL, int C, int DIR, bool LOAD_J_FAST, int R>
FFT_HD void tile_pf_body (const TensorMap5* imap, const TileArgs& a, const TileTmaCoords& ic, int distance)
{
    if (threadIdx.x == 0)
    {
        const long long t = (long long) blockIdx.x + distance;
        if (t < (long long) a.ntiles * a.batch)
        {
            const int bx = (int) (t / a.ntiles);
            const int g = (int) (t - (long long) bx * a.ntiles);
            const int ghi = g / a.gdiv, glo = g - ghi * a.gdiv;
            for (int it = 0; it < ic.iters; ++it)
            {
                int c[5];
#pragma unroll
                for (int d = 0; d < 5; ++d)
                    c[d] = ghi * ic.per_hi[d] + glo * ic.per_lo[d];
                c[ic.batch_dim] += bx;
                c[ic.iter_dim] += it * ic.iter_step;
                tma_prefetch_5d (imap, c[0], c[1], c[2], c[3], c[4]);
            }
        }
    }
    tile_body<LOGL, C, DIR, LOAD_J_FAST, 0, R> (a);
}
#ifndef CHOWDSP_EMU
template <int LOGL, int C, int DIR, bool LOAD_J_FAST, int R = 16>
__global__ void __launch_bounds__ (TileLaunch<LOGL, C, R>::THREADS, TileLaunch<LOGL, C, R>::MIN_BLOCKS) tile_fft_pf_kernel (CFB_TMAP5_PARAM imap, const TileArgs a, const TileTmaCoords ic, const int distance)
{
    tile_pf_body<LOGL, C, DIR, LOAD_J_FAST, R> (&imap, a, ic, distance);
}
#endif

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


#ifndef CHOWDSP_EMU
// ---------------------------------------------------------------------------------------------
// Cross-rank barrier of the distributed transform, in the stream, without a host round trip or a collective library:
// every rank owns a flag row flags[which][world] in peer-mapped memory.  One CTA per rank; lane g publishes this rank's
// step counter into rank g's row (release at system scope, after a system fence that orders the phase kernel's peer
// stores -- complete at kernel boundary -- before it) and then waits until rank g's counter has arrived in its own row.
// A rank that never arrives would hang the GPU, so the wait gives up after `timeout_ns` and raises *status instead.
// ---------------------------------------------------------------------------------------------
struct DistBarrierArgs
{
    unsigned long long* peer_flags[8]; // rank g's flag block (2 rows of 8 counters), mapped into this process
    unsigned long long* own_flags;
    int* status;                       // set to 1 on timeout
    int rank, world, which;
    unsigned long long step;
    unsigned long long timeout_ns;
};
template <int UNUSED = 0> // a template only so that every translation unit including this header may hold a copy
__global__ void __launch_bounds__ (32) dist_barrier_kernel (const DistBarrierArgs a)
{
    const int g = (int) threadIdx.x;
    if (g >= a.world)
        return;
    __threadfence_system();
    unsigned long long* dst = a.peer_flags[g] + a.which * 8 + a.rank;
    asm volatile ("st.release.sys.global.u64 [%0], %1;" ::"l"(dst), "l"(a.step) : "memory");
    const unsigned long long* src = a.own_flags + a.which * 8 + g;
    unsigned long long t0, now, seen;
    asm volatile ("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;)
    {
        asm volatile ("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(src) : "memory");
        if (seen >= a.step)
            break;
        asm volatile ("mov.u64 %0, %%globaltimer;" : "=l"(now));
        if (now - t0 > a.timeout_ns)
        {
            *a.status = 1;
            break;
        }
        __nanosleep (200);
    }
    __threadfence_system();
}
#endif

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
