// One-pass transforms of 2^15 .. 2^17 complex points on a THREAD-BLOCK CLUSTER (sm_90+ clusters, distributed shared memory,
// tensor-map TMA): the sizes between "one CTA holds the transform" (<= 2^14, fft_kernels.cuh / pipe_kernels.cuh) and the
// multi-pass tile path (large_kernels.cuh).  A cluster of G = N / 8192 CTAs owns one transform; every input element is read
// from HBM once and every output element written once -- the reference does the same job with one in-cache FFTPACK pass
// sequence over one work buffer (/root/reference/simd/chowdsp_fft_impl_avx.cpp:430-490, :1848-1935).
//
// Index split  n = a + 16 g + 16 G c   (a < 16: position inside a 128-byte line, g < G: CTA rank, c < 512):
//   1. CTA g loads the lines  l = g (mod G)  of the transform with ONE tensor-map TMA descriptor (4-D view
//      [batch][c][g][32 floats], box 256 lines = 32 KB per copy; SASS UTMALDG) into a landing buffer [c][a]
//   2. 16 local 512-point FFTs over c (one per a; the C = 16 "adjacent transforms" geometry of the tile kernels:
//      thread = (a, j), 32 points per thread, 32 x 16 Stockham with one shared-memory exchange)            -> index kc
//   3. twiddle W_(512 G)^(g kc), then the radix-G butterfly ACROSS THE CTAs: an all-to-all through distributed shared
//      memory -- every thread stores its 32/G values for destination p straight into CTA p's buffer
//      (st.shared::cluster), two cluster barriers per transform                                                -> k' = kc + 512 kb
//   4. twiddle W_N^(a k'), transposition through shared memory so that one thread holds all 16 a of a k',
//      radix-16 over a, coalesced stores  X[k' + (N/16) ka]  (256-byte runs per warp)
// Interleaving the CTAs at LINE granularity on the input side and at run granularity on the output side is what makes
// a single cross-CTA exchange enough (a contiguous-chunk split would need two).
// Two CTAs of different clusters are resident per SM (256 threads, <= 128 registers, 74 KB of shared memory each), so the
// load / exchange / store phases of neighbouring transforms overlap as they do in fft_kernel.
#pragma once
#include "fft_kernels.cuh"
#include "large_kernels.cuh" // big_twiddle tables, tile_region_stride, unord_pair_offset
#include "pipe_kernels.cuh"  // mbarrier helpers

namespace cfb
{
// ---------------------------------------------------------------------------------------------
// tensor map: the real CUtensorMap on the device, a plain description in the CPU emulator
// ---------------------------------------------------------------------------------------------
#ifdef CHOWDSP_EMU
struct TensorMap4
{
    const char* base;
    unsigned long long dim[4];    // elements (floats) per dimension, innermost first
    unsigned long long stride[3]; // bytes between consecutive indices of dimensions 1..3
    unsigned box[4];
};
#define CFB_TMAP_PARAM const TensorMap4
#else
struct alignas (64) TensorMap4
{
    unsigned long long opaque[16]; // CUtensorMap
};
#define CFB_TMAP_PARAM const __grid_constant__ TensorMap4
#endif

// cp.async.bulk.tensor.4d global -> shared, completion on an mbarrier of this CTA
FFT_HD void tma_load_4d (void* dst, const TensorMap4* map, int c0, int c1, int c2, int c3, unsigned long long* bar)
{
#ifdef CHOWDSP_EMU
    char* d = static_cast<char*> (dst);
    unsigned long long bytes = 0;
    for (unsigned i3 = 0; i3 < map->box[3]; ++i3)
        for (unsigned i2 = 0; i2 < map->box[2]; ++i2)
            for (unsigned i1 = 0; i1 < map->box[1]; ++i1)
            {
                const char* src = map->base + (unsigned long long) (c3 + i3) * map->stride[2] + (unsigned long long) (c2 + i2) * map->stride[1]
                                  + (unsigned long long) (c1 + i1) * map->stride[0] + (unsigned long long) c0 * 4;
                std::memcpy (d + bytes, src, (size_t) map->box[0] * 4);
                bytes += (unsigned long long) map->box[0] * 4;
            }
    __atomic_fetch_add (bar, bytes, __ATOMIC_RELEASE);
#else
    asm volatile ("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                  ::"r"(smem_addr (dst)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_addr (bar)) : "memory");
#endif
}

// cp.async.bulk.prefetch.tensor: ask L2 for a box ahead of time (no shared-memory destination, no completion)
FFT_HD void tma_prefetch_4d (const TensorMap4* map, int c0, int c1, int c2, int c3)
{
#ifndef CHOWDSP_EMU
    asm volatile ("cp.async.bulk.prefetch.tensor.4d.L2.global [%0, {%1, %2, %3, %4}];" ::"l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
#else
    (void) map; (void) c0; (void) c1; (void) c2; (void) c3;
#endif
}

// ---------------------------------------------------------------------------------------------
// cluster primitives
// ---------------------------------------------------------------------------------------------
FFT_HD int cluster_rank()
{
#ifdef CHOWDSP_EMU
    return emu::ctx.cluster_rank;
#else
    unsigned r;
    asm volatile ("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return (int) r;
#endif
}
FFT_HD int cluster_id_x() // index of this cluster in the grid
{
#ifdef CHOWDSP_EMU
    return emu::ctx.cluster_id;
#else
    unsigned r;
    asm volatile ("mov.u32 %0, %%clusterid.x;" : "=r"(r));
    return (int) r;
#endif
}
FFT_HD int cluster_count_x()
{
#ifdef CHOWDSP_EMU
    return emu::ctx.cluster_count;
#else
    unsigned r;
    asm volatile ("mov.u32 %0, %%nclusterid.x;" : "=r"(r));
    return (int) r;
#endif
}
// split cluster barrier: arrive (release) ... wait (acquire).  In the emulator the whole barrier happens at the wait.
FFT_HD void cluster_arrive()
{
#ifndef CHOWDSP_EMU
    asm volatile ("barrier.cluster.arrive.release.aligned;" ::: "memory");
#endif
}
FFT_HD void cluster_wait()
{
#ifdef CHOWDSP_EMU
    emu::ctx.cluster_bar->arrive_and_wait();
#else
    asm volatile ("barrier.cluster.wait.acquire.aligned;" ::: "memory");
#endif
}
// "this thread is done reading": no memory ordering needed (a release fence here would first drain the thread's outstanding
// global stores -- ncu showed it as the top stall of the first version, profiles/r02_cluster_kernel.txt)
FFT_HD void cluster_arrive_relaxed()
{
#ifndef CHOWDSP_EMU
    asm volatile ("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
#endif
}
// Asynchronous store into the shared memory of CTA `rank` of this cluster, at the position `local` has in this CTA's own
// window; the 8 bytes are counted on the mbarrier that sits at `bar`'s position in the DESTINATION CTA (st.async,
// complete_tx): the receiver waits on its own barrier, no cluster barrier and no fence on the sender's side.
FFT_HD void dsmem_store_async (float2* local, unsigned long long* bar, int rank, float2 v)
{
#ifdef CHOWDSP_EMU
    char* peer = emu::ctx.cluster_smem[rank];
    *reinterpret_cast<float2*> (peer + (reinterpret_cast<char*> (local) - emu::ctx.smem)) = v;
    __atomic_fetch_add (reinterpret_cast<unsigned long long*> (peer + (reinterpret_cast<char*> (bar) - emu::ctx.smem)), 8ull, __ATOMIC_RELEASE);
#else
    unsigned remote, rbar;
    asm volatile ("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_addr (local)), "r"(rank));
    asm volatile ("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rbar) : "r"(smem_addr (bar)), "r"(rank));
    asm volatile ("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f32 [%0], {%1, %2}, [%3];" ::"r"(remote), "f"(v.x), "f"(v.y), "r"(rbar) : "memory");
#endif
}
// cp.async.bulk.tensor.3d shared -> global (bulk group), then wait until the shared-memory source has been read
FFT_HD void tma_store_3d_and_release (const void* src, const TensorMap4* map, int c0, int c1, int c2)
{
#ifdef CHOWDSP_EMU
    const char* sp = static_cast<const char*> (src);
    unsigned long long bytes = 0;
    for (unsigned i2 = 0; i2 < map->box[2]; ++i2)
        for (unsigned i1 = 0; i1 < map->box[1]; ++i1)
        {
            char* dst = const_cast<char*> (map->base) + (unsigned long long) (c2 + i2) * map->stride[1] + (unsigned long long) (c1 + i1) * map->stride[0] + (unsigned long long) c0 * 4;
            std::memcpy (dst, sp + bytes, (size_t) map->box[0] * 4);
            bytes += (unsigned long long) map->box[0] * 4;
        }
#else
    asm volatile ("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_addr (src)) : "memory");
    asm volatile ("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile ("cp.async.bulk.wait_group.read 0;" ::: "memory");
#endif
}
FFT_HD void tma_store_drain()
{
#ifndef CHOWDSP_EMU
    asm volatile ("cp.async.bulk.wait_group 0;" ::: "memory");
#endif
}

// ---------------------------------------------------------------------------------------------
// geometry
// ---------------------------------------------------------------------------------------------
template <int LOGG>
struct ClusterGeo
{
    static constexpr int G = 1 << LOGG;          // CTAs per cluster
    static constexpr int LOGN = 13 + LOGG;       // complex points per transform
    static constexpr int C = 16, LC = 512, R = 32;
    using GL = Geo<9, 32>;                        // the local 512-point transforms
    static constexpr int T = GL::T;               // 16 threads per local transform
    static constexpr int THREADS = T * C;         // 256
    static constexpr int PER = R / G;             // registers per destination CTA in the exchange
    static constexpr int RUN = 16 * PER;          // contiguous output bins per (ka, kb) of one CTA
    static constexpr int RS = tile_region_stride (GL::SMEM_F2, C);
    static constexpr int LAND_F2 = LC * C;        // 8192: landing image, received image and store image
    static constexpr int XCH_F2 = C * RS;
    static constexpr int TR_PITCH = C + 1;        // transposition image [k' local][a], one pad slot per row
    static constexpr int TR_F2 = LC * TR_PITCH;
    static constexpr int BUF_F2 = TR_F2 > XCH_F2 ? (TR_F2 > LAND_F2 ? TR_F2 : LAND_F2) : (XCH_F2 > LAND_F2 ? XCH_F2 : LAND_F2);
    // small tables behind the work buffer (float2 units): stage-1 twiddle rows of the local transforms (15 x 16), the two
    // levels of the inner twiddle (4 + 8), the two levels of the outer twiddle (PER x 16 and G x 16)
    static constexpr int TW1_OFFSET = BUF_F2 * 8;
    static constexpr int TWIN_OFFSET = TW1_OFFSET + 15 * 16 * 8;
    static constexpr int TWU_OFFSET = TWIN_OFFSET + 12 * 8;
    static constexpr int TWV_OFFSET = TWU_OFFSET + PER * C * 8;
    static constexpr int BAR_OFFSET = TWV_OFFSET + G * C * 8;   // two mbarriers: input tile, exchange
    static constexpr int SMEM_BYTES = BAR_OFFSET + 16;
    static constexpr unsigned TILE_BYTES = LAND_F2 * 8;            // per CTA per transform
    static constexpr int TMA_ROWS = 256;                            // lines per tensor copy (box limit)
    static_assert (G >= 2 && G <= 16 && R % G == 0, "2 .. 16 CTAs per cluster");
    static_assert (GL::S == 2 && GL::T == 16 && GL::RLAST == 16, "written for 32 x 16 local transforms");
    static_assert (BAR_OFFSET % 8 == 0, "mbarrier alignment");
};

struct ClusterArgs
{
    float* out;
    long long out_stride; // floats between consecutive transforms
    int batch;
    int logW;             // 0: natural order output; 3: the reference's 8-lane unordered layout (forward transforms)
    const float2* tw;     // stage twiddles of Geo<9, 32>
    const float2* tw_lo;  // two-level table of W_N (large_kernels.cuh: fill_big_twiddles)
    const float2* tw_hi;
    int tw_lobits;
    int l2_prefetch;      // 1: tensor-map L2 prefetch of the next transform at the start of every iteration
};

FFT_HD float2 wn_pow (const ClusterArgs& a, unsigned e) // forward twiddle W_N^e; callers conjugate through cmul_dir<DIR>
{
    const float2 lo = __ldg (a.tw_lo + (e & ((1u << a.tw_lobits) - 1u)));
    const float2 hi = __ldg (a.tw_hi + (e >> a.tw_lobits));
    return cmul_dir<-1> (lo, hi);
}

// last stage of the local 512-point transforms (radix 16, two butterflies per thread) with its twiddle rows in shared
// memory: every acquire at cluster scope invalidates L1, so as global loads they would come from L2 once per transform
template <class GL, int DIR>
FFT_HD void cluster_stage1 (float2 (&v)[32], int j, const float2* tw1s)
{
    constexpr int r = 16, SUB = 2, T = 16;
    const float2* t = tw1s + j;
#pragma unroll
    for (int q = 1; q < r; ++q)
    {
        const float2 wq = lds2 (t + (q - 1) * T);
#pragma unroll
        for (int u = 0; u < SUB; ++u)
        {
            const float2 x = cmul_dir<DIR> (v[u + q * SUB], wq);
            v[u + q * SUB] = mul_w32_rt<DIR> (x, u * q);
        }
    }
#pragma unroll
    for (int u = 0; u < SUB; ++u)
        RegFft<r, DIR, SUB>::run (&v[u]);
}

// TMA_OUT: natural-order output through ONE tensor-map store per transform and CTA (the result is staged in the work
// buffer as the [ka G + kb][RUN] image of the CTA's part of the spectrum); otherwise per-thread 64-bit stores (the
// unordered layout's pair addressing)
template <int LOGG, int DIR, int LOGW>
FFT_HD void cluster_body (const TensorMap4* tmap, const TensorMap4* omap, const ClusterArgs& a)
{
    using CG = ClusterGeo<LOGG>;
    using GL = typename CG::GL;
    constexpr int G = CG::G, C = CG::C, R = CG::R, PER = CG::PER, RS = CG::RS, LOGN = CG::LOGN, RUN = CG::RUN;
    constexpr bool TMA_OUT = LOGW == 0;
    constexpr unsigned NMASK = (1u << LOGN) - 1u;
    FFT_DYN_SMEM (char, smem);
    float2* buf = reinterpret_cast<float2*> (smem);
    float2* sTw1 = reinterpret_cast<float2*> (smem + CG::TW1_OFFSET);
    float2* sTwIn = reinterpret_cast<float2*> (smem + CG::TWIN_OFFSET);  // [0..4): W^(256 g ml), [4..12): W^(1024 g mh)
    float2* sTwU = reinterpret_cast<float2*> (smem + CG::TWU_OFFSET);    // [ml][a] = W^(a 16 (g PER + ml))
    float2* sTwV = reinterpret_cast<float2*> (smem + CG::TWV_OFFSET);    // [kb][a] = W^(a 512 kb)
    unsigned long long* bar_in = reinterpret_cast<unsigned long long*> (smem + CG::BAR_OFFSET);
    unsigned long long* bar_x = bar_in + 1;

    const int tid = (int) threadIdx.x;
    const int lt = tid % C, j = tid / C; // column a = lt, thread j of that column's 512-point transform
    const int g = cluster_rank();
    const int nclusters = cluster_count_x();

    // ---- per-CTA tables (this CTA is exchange destination p = g) ----
    if (tid < 15 * 16)
        sts2 (sTw1 + tid, __ldg (a.tw + GL::tw_off (1) + tid));
    else
        smem_skip();
    if (tid < 12)
        sts2 (sTwIn + tid, wn_pow (a, (unsigned) (tid < 4 ? 256 * g * tid : 1024 * g * (tid - 4)) & NMASK));
    else
        smem_skip();
    for (int i = tid; i < PER * C; i += CG::THREADS)
        sts2 (sTwU + i, wn_pow (a, (unsigned) ((i % C) * 16 * (g * PER + i / C)) & NMASK));
    for (int i = tid; i < G * C; i += CG::THREADS)
        sts2 (sTwV + i, wn_pow (a, (unsigned) ((i % C) * 512 * (i / C)) & NMASK));
    const float2 win_j = wn_pow (a, (unsigned) (16 * g * j) & NMASK); // W_N^(16 g j)
    const float2 wout_j = wn_pow (a, (unsigned) (lt * j));             // W_N^(a j)
    if (tid == 0)
    {
        mbar_init (bar_in);
        mbar_init (bar_x);
    }
    __syncthreads();
    cluster_arrive(); // every CTA's barriers and tables exist before any peer touches this CTA's shared memory
    cluster_wait();

    auto fetch = [&] (int b)
    {
        mbar_expect (bar_in, CG::TILE_BYTES);
#pragma unroll
        for (int h = 0; h < CG::LC / CG::TMA_ROWS; ++h)
            tma_load_4d (buf + h * CG::TMA_ROWS * C, tmap, 0, g, h * CG::TMA_ROWS, b, bar_in);
    };
    int b = cluster_id_x();
    if (tid == 0 && b < a.batch)
        fetch (b);

    for (unsigned it = 0; b < a.batch; b += nclusters, ++it)
    {
        float2 v[R];
        // The work buffer doubles as the landing buffer, so the NEXT transform's copy can only start at the end of this
        // iteration: ask L2 for it now (ncu on the first version: 23 % of all stall samples sat on the input barrier), the
        // copy itself is then an L2 hit
        if (tid == 0 && a.l2_prefetch != 0 && b + nclusters < a.batch)
        {
#pragma unroll
            for (int h = 0; h < CG::LC / CG::TMA_ROWS; ++h)
                tma_prefetch_4d (tmap, 0, g, h * CG::TMA_ROWS, b + nclusters);
        }
        mbar_wait (bar_in, it, CG::TILE_BYTES);
        // ---- 1/2: stage-0 registers from the landing buffer, v[m] = x[a + 16 g + 16 G (j + 16 m)] ----
        {
            const float2* lj = buf + tid;
#pragma unroll
            for (int m = 0; m < R; ++m)
                v[m] = lds2 (lj + m * CG::THREADS);
        }
        float2* sB = buf + lt * RS;
        stage_compute<GL, DIR, 0> (v, j, a.tw);
        __syncthreads(); // the landing image has been consumed
        stage_scatter<GL, 0> (v, j, sB);
        __syncthreads();
        gather_natural<GL, 0, R> (v, j, sB);
        cluster_arrive_relaxed(); // this thread no longer reads `buf`: peers may start filling it once everybody has arrived
        if (tid == 0)
            mbar_expect (bar_x, CG::TILE_BYTES); // the exchange delivers one full image (own share included)
        cluster_stage1<GL, DIR> (v, j, sTw1);
        // ---- 3: inner twiddle W_(512 G)^(g kc) = W^(16 g j) * W^(256 g (m mod 4)) * W^(1024 g (m div 4)), then the all-to-all ----
        if (g != 0)
        {
            float2 wa[4];
            wa[0] = win_j;
#pragma unroll
            for (int q = 1; q < 4; ++q)
                wa[q] = cmul_dir<-1> (win_j, lds2 (sTwIn + q));
#pragma unroll
            for (int mh = 0; mh < 8; ++mh)
            {
                const float2 wb = lds2 (sTwIn + 4 + mh);
#pragma unroll
                for (int q = 0; q < 4; ++q)
                {
                    const float2 w = mh == 0 ? wa[q] : cmul_dir<-1> (wa[q], wb);
                    v[4 * mh + q] = cmul_dir<DIR> (v[4 * mh + q], w);
                }
            }
        }
        else
        {
#pragma unroll
            for (int q = 0; q < 3 + 8; ++q)
                smem_skip();
        }
        cluster_wait();
        // value m goes to CTA p = m / PER, slot (g PER + m % PER) of its buffer, position tid inside the slot
#pragma unroll
        for (int m = 0; m < R; ++m)
            dsmem_store_async (buf + (g * PER + m % PER) * CG::THREADS + tid, bar_x, m / PER, v[m]);
        mbar_wait (bar_x, it, CG::TILE_BYTES);
        // v[g' PER + ml] = value of CTA g' for kc = j + 16 (g PER + ml); radix-G butterfly over g' -> kb
        {
            const float2* rj = buf + tid;
#pragma unroll
            for (int i = 0; i < R; ++i)
                v[i] = lds2 (rj + i * CG::THREADS);
        }
#pragma unroll
        for (int ml = 0; ml < PER; ++ml)
            RegFft<G, DIR, PER>::run (&v[ml]);
        // ---- 4: outer twiddle W_N^(a k') = W^(a j) * W^(a 16 (g PER + ml)) * W^(a 512 kb), k' = j + 16 (g PER + ml) + 512 kb ----
        {
            float2 wu[PER];
#pragma unroll
            for (int ml = 0; ml < PER; ++ml)
                wu[ml] = cmul_dir<-1> (wout_j, lds2 (sTwU + ml * C + lt));
#pragma unroll
            for (int kb = 0; kb < G; ++kb)
            {
                const float2 wv = lds2 (sTwV + kb * C + lt);
#pragma unroll
                for (int ml = 0; ml < PER; ++ml)
                {
                    const float2 w = kb == 0 ? wu[ml] : cmul_dir<-1> (wu[ml], wv);
                    v[kb * PER + ml] = cmul_dir<DIR> (v[kb * PER + ml], w);
                }
            }
        }
        __syncthreads(); // the received image has been consumed
        // transposition image: row k'_local = i 16 + j, column a
        {
            float2* tj = buf + j * CG::TR_PITCH + lt;
#pragma unroll
            for (int i = 0; i < R; ++i)
                sts2 (tj + i * 16 * CG::TR_PITCH, v[i]);
        }
        __syncthreads();
        float2 x[2][16];
#pragma unroll
        for (int h = 0; h < 2; ++h)
        {
            const float2* row = buf + (tid + h * CG::THREADS) * CG::TR_PITCH;
#pragma unroll
            for (int aa = 0; aa < 16; ++aa)
                x[h][aa] = lds2 (row + aa);
        }
        if constexpr (TMA_OUT)
        {
            // store image [ka G + kb][RUN]: this CTA's part of the spectrum, bins 16 (g PER) + r + 512 (kb + G ka), r < RUN
            __syncthreads(); // the transposition image has been consumed
#pragma unroll
            for (int h = 0; h < 2; ++h)
            {
                RegFft<16, DIR, 1>::run (x[h]);
                const int kl = tid + h * CG::THREADS; // k'_local = i 16 + jj, i = kb PER + ml
                const int i = kl >> 4, jj = kl & 15;
                const int kb = i / PER, ml = i % PER;
                float2* img = buf + kb * RUN + ml * 16 + jj;
#pragma unroll
                for (int ka = 0; ka < 16; ++ka)
                    sts2 (img + ka * G * RUN, x[h][ka]);
            }
            fence_proxy_async(); // generic-proxy writes of the image are ordered before the async-proxy (TMA) reads
            __syncthreads();
            if (tid == 0)
            {
                tma_store_3d_and_release (buf, omap, 2 * RUN * g, 0, b); // returns once the image has been read out of `buf`
                if (b + nclusters < a.batch)
                    fetch (b + nclusters);
            }
        }
        else
        {
            fence_proxy_async(); // generic-proxy accesses to `buf` are ordered before the next tile's async-proxy (TMA) writes
            __syncthreads();     // `buf` is free: fetch the next transform while this one is finished and stored
            if (tid == 0 && b + nclusters < a.batch)
                fetch (b + nclusters);
            float* __restrict__ ob = a.out + (long long) b * a.out_stride;
#pragma unroll
            for (int h = 0; h < 2; ++h)
            {
                RegFft<16, DIR, 1>::run (x[h]);
                const int kl = tid + h * CG::THREADS;
                const int i = kl >> 4, jj = kl & 15;
                const int kb = i / PER, ml = i % PER;
                const int kp = jj + 16 * (g * PER + ml) + 512 * kb; // k'
                // unordered output: threads with k' and k' ^ 1 (adjacent lanes) swap one float: the even one stores (re, re'),
                // the odd one (im, im') -- contiguous 8-byte pairs of the lane layout (as fft_core's UDIRECT)
                const int odd = jj & 1;
#pragma unroll
                for (int ka = 0; ka < 16; ++ka)
                {
                    const float recv = shfl1 (odd ? x[h][ka].x : x[h][ka].y, (tid ^ 1) & 31, 32);
                    const float2 o = odd ? make_float2 (recv, x[h][ka].y) : make_float2 (x[h][ka].x, recv);
                    *reinterpret_cast<float2*> (ob + unord_pair_offset ((long long) kp + ((long long) ka << (LOGN - 4)), LOGN, LOGW)) = o;
                }
            }
        }
    }
    if (TMA_OUT && tid == 0)
        tma_store_drain();
    // nobody leaves while a peer may still store into its shared memory
    cluster_arrive();
    cluster_wait();
}

#ifndef CHOWDSP_EMU
template <int LOGG, int DIR, int LOGW>
__global__ void __launch_bounds__ (ClusterGeo<LOGG>::THREADS, 2) cluster_fft_kernel (CFB_TMAP_PARAM tmap, CFB_TMAP_PARAM omap, const ClusterArgs a)
{
    cluster_body<LOGG, DIR, LOGW> (&tmap, &omap, a);
}
#else
template <int LOGG, int DIR, int LOGW>
void cluster_fft_kernel (CFB_TMAP_PARAM tmap, CFB_TMAP_PARAM omap, const ClusterArgs a)
{
    cluster_body<LOGG, DIR, LOGW> (&tmap, &omap, a);
}
#endif
} // namespace cfb
