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
// store into the shared memory of CTA `rank` of this cluster, at the position `local` has in this CTA's own window
FFT_HD void dsmem_store (float2* local, int rank, float2 v)
{
#ifdef CHOWDSP_EMU
    *reinterpret_cast<float2*> (emu::ctx.cluster_smem[rank] + (reinterpret_cast<char*> (local) - emu::ctx.smem)) = v;
#else
    unsigned remote;
    asm volatile ("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_addr (local)), "r"(rank));
    asm volatile ("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(remote), "f"(v.x), "f"(v.y) : "memory");
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
    static constexpr int RS = tile_region_stride (GL::SMEM_F2, C);
    static constexpr int LAND_F2 = LC * C;        // 8192
    static constexpr int XCH_F2 = C * RS;
    static constexpr int TR_PITCH = C + 1;        // transposition image [k' local][a], one pad slot per row
    static constexpr int TR_F2 = LC * TR_PITCH;
    static constexpr int BUF_F2 = TR_F2 > XCH_F2 ? (TR_F2 > LAND_F2 ? TR_F2 : LAND_F2) : (XCH_F2 > LAND_F2 ? XCH_F2 : LAND_F2);
    static constexpr int TWIN_OFFSET = BUF_F2 * 8;                 // 32 inner-twiddle steps
    static constexpr int TWOUT_OFFSET = TWIN_OFFSET + R * 8;       // 32 x 16 outer-twiddle steps
    static constexpr int BAR_OFFSET = TWOUT_OFFSET + R * C * 8;
    static constexpr int SMEM_BYTES = BAR_OFFSET + 16;
    static constexpr unsigned TILE_BYTES = LAND_F2 * 8;            // per CTA per transform
    static constexpr int TMA_ROWS = 256;                            // lines per tensor copy (box limit)
    static_assert (G >= 2 && G <= 16 && R % G == 0, "2 .. 16 CTAs per cluster");
    static_assert (GL::S == 2 && GL::T == 16, "written for 32 x 16 local transforms");
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
};

template <int DIR>
FFT_HD float2 wn_pow (const ClusterArgs& a, unsigned e)
{
    const float2 lo = __ldg (a.tw_lo + (e & ((1u << a.tw_lobits) - 1u)));
    const float2 hi = __ldg (a.tw_hi + (e >> a.tw_lobits));
    return cmul_dir<-1> (lo, hi); // forward twiddle W_N^e; callers conjugate through cmul_dir<DIR>
}

template <int LOGG, int DIR, int LOGW>
FFT_HD void cluster_body (const TensorMap4* tmap, const ClusterArgs& a)
{
    using CG = ClusterGeo<LOGG>;
    using GL = typename CG::GL;
    constexpr int G = CG::G, C = CG::C, T = CG::T, R = CG::R, PER = CG::PER, RS = CG::RS, LOGN = CG::LOGN;
    constexpr unsigned NMASK = (1u << LOGN) - 1u;
    FFT_DYN_SMEM (char, smem);
    float2* buf = reinterpret_cast<float2*> (smem);
    float2* sTwIn = reinterpret_cast<float2*> (smem + CG::TWIN_OFFSET);
    float2* sTwOut = reinterpret_cast<float2*> (smem + CG::TWOUT_OFFSET);
    unsigned long long* bar = reinterpret_cast<unsigned long long*> (smem + CG::BAR_OFFSET);

    const int tid = (int) threadIdx.x;
    const int lt = tid % C, j = tid / C; // column a = lt, thread j of that column's 512-point transform
    const int g = cluster_rank();
    const int nclusters = cluster_count_x();

    // per-CTA twiddle steps (see steps 3 and 4 below); this CTA is exchange destination p = g
    //   sTwIn [m]       = W_N^(16 g * 16 m)                                   inner twiddle W_(512 G)^(g kc), kc = j + 16 m
    //   sTwOut[i][a]    = W_N^(a * (16 (g PER + ml) + 512 kb)), i = kb PER + ml   outer twiddle W_N^(a k')
    if (tid < R)
        sts2 (sTwIn + tid, wn_pow<DIR> (a, (unsigned) (16 * g * 16 * tid) & NMASK));
    else
        smem_skip();
    for (int i = tid; i - tid < R * C; i += CG::THREADS)
    {
        const int reg = i / C, aa = i % C;
        const int kb = reg / PER, ml = reg % PER;
        sts2 (sTwOut + i, wn_pow<DIR> (a, (unsigned) (aa * (16 * (g * PER + ml) + 512 * kb)) & NMASK));
    }
    const float2 win_j = wn_pow<DIR> (a, (unsigned) (16 * g * j) & NMASK);  // W_N^(16 g j)
    const float2 wout_j = wn_pow<DIR> (a, (unsigned) (lt * j));              // W_N^(a j)
    if (tid == 0)
        mbar_init (bar);
    __syncthreads();
    cluster_arrive(); // every CTA's barrier and tables exist before any peer touches this CTA's shared memory
    cluster_wait();

    auto fetch = [&] (int b)
    {
        mbar_expect (bar, CG::TILE_BYTES);
#pragma unroll
        for (int h = 0; h < CG::LC / CG::TMA_ROWS; ++h)
            tma_load_4d (buf + h * CG::TMA_ROWS * C, tmap, 0, g, h * CG::TMA_ROWS, b, bar);
    };
    int b = cluster_id_x();
    if (tid == 0 && b < a.batch)
        fetch (b);

    for (unsigned it = 0; b < a.batch; b += nclusters, ++it)
    {
        float2 v[R];
        mbar_wait (bar, it, CG::TILE_BYTES);
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
        cluster_arrive(); // this thread no longer reads `buf`: peers may start filling it (after everybody has arrived)
        stage_compute<GL, DIR, 1> (v, j, a.tw);
        // ---- 3: inner twiddle W_(512 G)^(g kc) = W_N^(16 g j) * W_N^(16 g 16 m), then the all-to-all ----
        if (g != 0)
        {
#pragma unroll
            for (int m = 0; m < R; ++m)
            {
                const float2 w = m == 0 ? win_j : cmul_dir<-1> (win_j, lds2 (sTwIn + m));
                v[m] = cmul_dir<DIR> (v[m], w);
            }
        }
        else
        {
#pragma unroll
            for (int m = 1; m < R; ++m)
                smem_skip();
        }
        cluster_wait();
        // value m goes to CTA p = m / PER, slot (g PER + m % PER) of its buffer, position tid inside the slot
#pragma unroll
        for (int m = 0; m < R; ++m)
            dsmem_store (buf + (g * PER + m % PER) * CG::THREADS + tid, m / PER, v[m]);
        cluster_arrive();
        cluster_wait();
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
        // ---- 4: outer twiddle W_N^(a k'), k' = j + 16 (g PER + ml) + 512 kb ----
#pragma unroll
        for (int i = 0; i < R; ++i)
        {
            const float2 w = cmul_dir<-1> (wout_j, lds2 (sTwOut + i * C + lt));
            v[i] = cmul_dir<DIR> (v[i], w);
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
        fence_proxy_async(); // generic-proxy accesses to `buf` are ordered before the next tile's async-proxy (TMA) writes
        __syncthreads();     // `buf` is free: fetch the next transform while this one is finished and stored
        if (tid == 0 && b + nclusters < a.batch)
            fetch (b + nclusters);
        float* __restrict__ ob = a.out + (long long) b * a.out_stride;
#pragma unroll
        for (int h = 0; h < 2; ++h)
        {
            RegFft<16, DIR, 1>::run (x[h]);
            const int kl = tid + h * CG::THREADS; // k'_local = i 16 + jj
            const int i = kl >> 4, jj = kl & 15;
            const int kb = i / PER, ml = i % PER;
            const int kp = jj + 16 * (g * PER + ml) + 512 * kb; // k'
            if constexpr (LOGW == 0)
            {
                float2* __restrict__ o2 = reinterpret_cast<float2*> (ob) + kp;
#pragma unroll
                for (int ka = 0; ka < 16; ++ka)
                    o2[(long long) ka << (LOGN - 4)] = x[h][ka];
            }
            else
            {
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
    // nobody leaves while a peer may still store into its shared memory
    cluster_arrive();
    cluster_wait();
}

#ifndef CHOWDSP_EMU
template <int LOGG, int DIR, int LOGW>
__global__ void __launch_bounds__ (ClusterGeo<LOGG>::THREADS, 2) cluster_fft_kernel (CFB_TMAP_PARAM tmap, const ClusterArgs a)
{
    cluster_body<LOGG, DIR, LOGW> (&tmap, a);
}
#else
template <int LOGG, int DIR, int LOGW>
void cluster_fft_kernel (CFB_TMAP_PARAM tmap, const ClusterArgs a)
{
    cluster_body<LOGG, DIR, LOGW> (&tmap, a);
}
#endif
} // namespace cfb
