// TEST INFRASTRUCTURE ONLY -- host shim that lets chowdsp_fft_b200/csrc/*.cuh compile with plain g++.
//
// There is no GPU in the build container, so index maps, twiddle tables, layouts and shared-memory
// bank behaviour of the CUDA kernels are pre-validated here: every CUDA thread of a block becomes one
// OS thread, __syncthreads() is a std::barrier, dynamic shared memory is a heap block.  The kernels'
// source is compiled UNCHANGED (-DCHOWDSP_EMU).  This is a checker for the kernel source, not a CPU
// implementation of the product: nothing under chowdsp_fft_b200/ builds, links or loads it, and the
// shipped library has no CPU path.
#pragma once
#include <algorithm>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <thread>
#include <type_traits>
#include <vector>

struct float2 { float x, y; };
struct alignas (16) float4 { float x, y, z, w; };
static inline float2 make_float2 (float x, float y) { return float2 { x, y }; }
static inline float4 make_float4 (float x, float y, float z, float w) { return float4 { x, y, z, w }; }
struct dim3
{
    unsigned x = 1, y = 1, z = 1;
    dim3 (unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x (x_), y (y_), z (z_) {}
};
typedef void* cudaStream_t;

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)

namespace emu
{
struct SmemOp { uint32_t off; uint8_t bytes; uint8_t is_store; };
struct ThreadCtx
{
    dim3 tIdx, bIdx, bDim, gDim;
    std::barrier<>* bar = nullptr;
    std::barrier<>* group_bar[3] = { nullptr, nullptr, nullptr }; // barriers of this thread's group of 32 / 64 / 128 threads
    char* smem = nullptr;
    std::vector<SmemOp>* log = nullptr; // per-thread smem op sequence (bank-conflict analysis)
    float* shfl_buf = nullptr;          // 32 floats of this thread's warp (warp shuffles)
    int lane = 0;
    // thread-block clusters (launch_cluster): rank of this CTA, index / count of clusters, every CTA's shared-memory
    // block (distributed shared memory) and the cluster-wide barrier
    int cluster_rank = 0, cluster_id = 0, cluster_count = 1;
    char* cluster_smem[16] = {};
    std::barrier<>* cluster_bar = nullptr;
};
inline thread_local ThreadCtx ctx;
inline void syncgroup (int n) { ctx.group_bar[n == 32 ? 0 : n == 64 ? 1 : 2]->arrive_and_wait(); } // __syncwarp / named barriers

// __shfl_sync (full mask) with a sub-group width: publish, barrier, read the source lane's value, barrier
inline float shfl (float v, int src, int width)
{
    ctx.shfl_buf[ctx.lane] = v;
    syncgroup (32);
    const float r = ctx.shfl_buf[(ctx.lane & ~(width - 1)) + (src & (width - 1))];
    syncgroup (32);
    return r;
}

struct ConflictStats
{
    long ops = 0;           // warp-level shared-memory instructions
    long wavefronts = 0;    // modelled wavefronts
    long ideal = 0;         // wavefronts with zero conflicts
    int worst = 0;          // worst wavefronts/ideal ratio seen (x100)
};
inline ConflictStats g_stats;
inline bool g_log_smem = false;

// wavefront model: an access of 32 lanes costs max over the 32 banks of the number of DISTINCT 4-byte
// words requested in that bank (same-word requests broadcast).
inline void analyse_block (std::vector<std::vector<SmemOp>>& logs)
{
    const size_t nthreads = logs.size();
    for (size_t w0 = 0; w0 < nthreads; w0 += 32)
    {
        const size_t w1 = std::min (nthreads, w0 + 32);
        const size_t nops = logs[w0].size();
        for (size_t t = w0; t < w1; ++t)
            if (logs[t].size() != nops)
            {
                std::fprintf (stderr, "emu: divergent smem op count inside a warp (%zu vs %zu)\n", logs[t].size(), nops);
                std::abort();
            }
        for (size_t i = 0; i < nops; ++i)
        {
            std::vector<uint32_t> words[32];
            int bytes = 0;
            for (size_t t = w0; t < w1; ++t)
            {
                const SmemOp& op = logs[t][i];
                if (op.bytes == 0)
                    continue; // predicated off
                bytes = std::max<int> (bytes, op.bytes);
                for (uint32_t b = 0; b < op.bytes; b += 4)
                {
                    const uint32_t word = (op.off + b) / 4;
                    auto& v = words[word % 32];
                    if (std::find (v.begin(), v.end(), word) == v.end())
                        v.push_back (word);
                }
            }
            if (bytes == 0)
                continue;
            int wf = 0;
            size_t total_words = 0;
            for (auto& v : words)
            {
                wf = std::max<int> (wf, (int) v.size());
                total_words += v.size();
            }
            const int ideal = std::max<int> (1, (int) ((total_words + 31) / 32));
            g_stats.ops++;
            g_stats.wavefronts += wf;
            g_stats.ideal += ideal;
            g_stats.worst = std::max (g_stats.worst, wf * 100 / ideal);
        }
    }
}

template <typename Kernel, typename... Args>
void launch (Kernel kernel, dim3 grid, dim3 block, size_t smem_bytes, Args... args)
{
    const unsigned nthreads = block.x * block.y * block.z;
    for (unsigned bx = 0; bx < grid.x; ++bx)
    {
        std::vector<char> smem (smem_bytes + 64);
        std::barrier<> bar ((std::ptrdiff_t) nthreads);
        std::vector<std::unique_ptr<std::barrier<>>> group_bars[3];
        for (int gi = 0; gi < 3; ++gi)
            for (unsigned w = 0, n = 32u << gi; w < (nthreads + n - 1) / n; ++w)
                group_bars[gi].emplace_back (new std::barrier<> ((std::ptrdiff_t) std::min (n, nthreads - n * w)));
        std::vector<std::vector<SmemOp>> logs (nthreads);
        std::vector<float> shfl_bufs ((nthreads + 31) / 32 * 32);
        std::vector<std::thread> pool;
        pool.reserve (nthreads);
        for (unsigned t = 0; t < nthreads; ++t)
            pool.emplace_back ([&, t]
                               {
                                   ctx.tIdx = dim3 (t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
                                   ctx.bIdx = dim3 (bx, 0, 0);
                                   ctx.bDim = block;
                                   ctx.gDim = grid;
                                   ctx.bar = &bar;
                                   for (int gi = 0; gi < 3; ++gi)
                                       ctx.group_bar[gi] = group_bars[gi][t / (32u << gi)].get();
                                   ctx.smem = smem.data();
                                   ctx.shfl_buf = shfl_bufs.data() + (t / 32) * 32;
                                   ctx.lane = (int) (t % 32);
                                   ctx.log = g_log_smem ? &logs[t] : nullptr;
                                   kernel (args...);
                               });
        for (auto& th : pool)
            th.join();
        if (g_log_smem)
            analyse_block (logs);
    }
}

// Cluster launch: the `cluster` CTAs of one cluster run CONCURRENTLY (they exchange data through each other's shared
// memory and meet at cluster barriers); clusters run one after the other.  grid.x = clusters * cluster.
template <typename Kernel, typename... Args>
void launch_cluster (Kernel kernel, dim3 grid, unsigned cluster, dim3 block, size_t smem_bytes, Args... args)
{
    const unsigned nthreads = block.x * block.y * block.z;
    const unsigned nclusters = grid.x / cluster;
    for (unsigned cid = 0; cid < nclusters; ++cid)
    {
        std::vector<std::vector<char>> smem (cluster, std::vector<char> (smem_bytes + 64));
        std::vector<std::unique_ptr<std::barrier<>>> bars;
        std::barrier<> cluster_bar ((std::ptrdiff_t) (nthreads * cluster));
        std::vector<std::vector<std::unique_ptr<std::barrier<>>>> group_bars (cluster * 3);
        std::vector<std::vector<std::vector<SmemOp>>> logs (cluster, std::vector<std::vector<SmemOp>> (nthreads));
        std::vector<std::vector<float>> shfl_bufs (cluster, std::vector<float> ((nthreads + 31) / 32 * 32));
        for (unsigned r = 0; r < cluster; ++r)
        {
            bars.emplace_back (new std::barrier<> ((std::ptrdiff_t) nthreads));
            for (int gi = 0; gi < 3; ++gi)
                for (unsigned w = 0, n = 32u << gi; w < (nthreads + n - 1) / n; ++w)
                    group_bars[r * 3 + gi].emplace_back (new std::barrier<> ((std::ptrdiff_t) std::min (n, nthreads - n * w)));
        }
        std::vector<std::thread> pool;
        pool.reserve ((size_t) nthreads * cluster);
        for (unsigned r = 0; r < cluster; ++r)
            for (unsigned t = 0; t < nthreads; ++t)
                pool.emplace_back ([&, r, t]
                                   {
                                       ctx.tIdx = dim3 (t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
                                       ctx.bIdx = dim3 (cid * cluster + r, 0, 0);
                                       ctx.bDim = block;
                                       ctx.gDim = grid;
                                       ctx.bar = bars[r].get();
                                       for (int gi = 0; gi < 3; ++gi)
                                           ctx.group_bar[gi] = group_bars[r * 3 + gi][t / (32u << gi)].get();
                                       ctx.smem = smem[r].data();
                                       ctx.shfl_buf = shfl_bufs[r].data() + (t / 32) * 32;
                                       ctx.lane = (int) (t % 32);
                                       ctx.log = g_log_smem ? &logs[r][t] : nullptr;
                                       ctx.cluster_rank = (int) r;
                                       ctx.cluster_id = (int) cid;
                                       ctx.cluster_count = (int) nclusters;
                                       for (unsigned q = 0; q < cluster; ++q)
                                           ctx.cluster_smem[q] = smem[q].data();
                                       ctx.cluster_bar = &cluster_bar;
                                       kernel (args...);
                                   });
        for (auto& th : pool)
            th.join();
        if (g_log_smem)
            for (unsigned r = 0; r < cluster; ++r)
                analyse_block (logs[r]);
    }
}
} // namespace emu

#define threadIdx (emu::ctx.tIdx)
#define blockIdx (emu::ctx.bIdx)
#define blockDim (emu::ctx.bDim)
#define gridDim (emu::ctx.gDim)
static inline void __syncthreads() { emu::ctx.bar->arrive_and_wait(); }
template <typename T>
static inline T __ldg (const T* p) { return *p; }
static inline float __fmaf_rn (float a, float b, float c) { return std::fmaf (a, b, c); }
