// Instantiations and launchers of the multi-pass (large transform) kernels.
#include <cstdint>

#include "dispatch.h"
#include "large_plan.h"

namespace cfb
{
namespace
{
template <int LOGL, int C, int DIR, bool JFAST>
cudaError_t launch_tile_one (const TileArgs& a, cudaStream_t stream)
{
    using TL = TileLaunch<LOGL, C>;
    auto kernel = tile_fft_kernel<LOGL, C, DIR, JFAST>;
    if (TL::SMEM_BYTES > 48 * 1024)
    {
        const cudaError_t e = cudaFuncSetAttribute (kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TL::SMEM_BYTES);
        if (e != cudaSuccess)
            return e;
    }
    kernel<<<(unsigned) a.ntiles * (unsigned) a.batch, TL::THREADS, TL::SMEM_BYTES, stream>>> (a);
    count_launch();
    return cudaGetLastError();
}
// persistent TMA-staged variant; cudaErrorInvalidConfiguration = does not apply (buffers do not fit / unaligned rows)
template <int LOGL, int C, int DIR, bool JFAST>
cudaError_t launch_tile_pipe_one (const TileArgs& a, cudaStream_t stream)
{
    using TP = TilePipeLaunch<LOGL, C>;
    if constexpr (! TP::FITS)
        return cudaErrorInvalidConfiguration;
    else
    {
        const auto even = [] (long long v) { return (v & 1) == 0; };
        if ((reinterpret_cast<uintptr_t> (a.in) & 15) != 0 || ! even (a.in_bstride) || ! even (a.in_g_hi) || ! even (a.in_g_lo)
            || (JFAST ? ! even (a.in_tstride) : (! even (a.in_estride) || (a.in_split_log < 31 && ! even (a.in_chunk_stride)))))
            return cudaErrorInvalidConfiguration;
        auto kernel = tile_pipe_kernel<LOGL, C, DIR, JFAST>;
        static thread_local int c_dev = -1, c_resident = 0;
        int dev = 0;
        cudaError_t e = cudaGetDevice (&dev);
        if (e != cudaSuccess)
            return e;
        if (dev != c_dev)
        {
            int sms = 0, per_sm = 0;
            if ((e = cudaFuncSetAttribute (kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TP::SMEM_BYTES)) != cudaSuccess
                || (e = cudaDeviceGetAttribute (&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess
                || (e = cudaOccupancyMaxActiveBlocksPerMultiprocessor (&per_sm, kernel, TP::THREADS, (size_t) TP::SMEM_BYTES)) != cudaSuccess)
                return e;
            if (per_sm < 1)
                return cudaErrorInvalidConfiguration;
            c_dev = dev;
            c_resident = sms * per_sm;
        }
        const long long tiles = (long long) a.ntiles * a.batch;
        if (tiles <= 0)
            return cudaSuccess;
        kernel<<<(unsigned) (tiles < c_resident ? tiles : c_resident), TP::THREADS, TP::SMEM_BYTES, stream>>> (a);
        count_launch();
        return cudaGetLastError();
    }
}
template <int LOGL, int C, int DIR, bool JFAST>
cudaError_t launch_tile_any (const TileArgs& a, cudaStream_t stream)
{
    if (tile_pipe_mode() != 0)
    {
        const cudaError_t e = launch_tile_pipe_one<LOGL, C, DIR, JFAST> (a, stream);
        if (e != cudaErrorInvalidConfiguration)
            return e;
        (void) cudaGetLastError();
    }
    return launch_tile_one<LOGL, C, DIR, JFAST> (a, stream);
}
template <int LOGL, int C>
cudaError_t launch_tile_lc (int dir, bool jfast, const TileArgs& a, cudaStream_t stream)
{
    if (dir < 0)
        return jfast ? launch_tile_any<LOGL, C, -1, true> (a, stream) : launch_tile_any<LOGL, C, -1, false> (a, stream);
    return jfast ? launch_tile_any<LOGL, C, +1, true> (a, stream) : launch_tile_any<LOGL, C, +1, false> (a, stream);
}
template <int LOGL>
cudaError_t launch_tile_l (int C, int dir, bool jfast, const TileArgs& a, cudaStream_t stream)
{
    if (C == 8)
        return launch_tile_lc<LOGL, 8> (dir, jfast, a, stream);
    if constexpr (LOGL <= 9)
        if (C == 16)
            return launch_tile_lc<LOGL, 16> (dir, jfast, a, stream);
    return cudaErrorInvalidValue;
}
} // namespace

cudaError_t launch_tile (int logL, int C, int dir, bool load_j_fast, const TileArgs& args, cudaStream_t stream)
{
    TileArgs a = args;
    a.pf_ahead = tile_pf_ahead();
    switch (logL)
    {
        case 6: return launch_tile_l<6> (C, dir, load_j_fast, a, stream);
        case 7: return launch_tile_l<7> (C, dir, load_j_fast, a, stream);
        case 8: return launch_tile_l<8> (C, dir, load_j_fast, a, stream);
        case 9: return launch_tile_l<9> (C, dir, load_j_fast, a, stream);
        case 10: return launch_tile_l<10> (C, dir, load_j_fast, a, stream);
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_real_pass (int dir, const RealPassArgs& a, int batch, cudaStream_t stream)
{
    const long long pairs = 1LL << (a.logM - 1);
    const dim3 grid ((unsigned) ((pairs + 255) / 256), (unsigned) batch);
    if (dir < 0)
        real_pass_kernel<-1><<<grid, 256, 0, stream>>> (a);
    else
        real_pass_kernel<+1><<<grid, 256, 0, stream>>> (a);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_complex_reorder (const float* in, float* out, long long in_bstride, long long out_bstride, int batch, int logN, int logW, bool to_unordered, cudaStream_t stream)
{
    const long long bins = 1LL << logN;
    const dim3 grid ((unsigned) ((bins + 255) / 256), (unsigned) batch);
    if (to_unordered)
        complex_reorder_kernel<true><<<grid, 256, 0, stream>>> (in, out, in_bstride, out_bstride, logN, logW);
    else
        complex_reorder_kernel<false><<<grid, 256, 0, stream>>> (in, out, in_bstride, out_bstride, logN, logW);
    count_launch();
    return cudaGetLastError();
}
} // namespace cfb
