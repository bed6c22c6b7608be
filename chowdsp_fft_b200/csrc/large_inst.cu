// Instantiations and launchers of the multi-pass (large transform) kernels.
#include <cstdint>

#include "dispatch.h"
#include "large_plan.h"
#include "tma_host.h"

namespace cfb
{
namespace
{
template <int LOGL, int C, int DIR, bool JFAST, int UIO, int R>
cudaError_t launch_tile_one (const TileArgs& a, cudaStream_t stream)
{
    using TL = TileLaunch<LOGL, C, R>;
    auto kernel = tile_fft_kernel<LOGL, C, DIR, JFAST, UIO, R>;
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
// uio: 0 natural order; 1 unordered input -- only ever the first (strided) pass of an inverse transform; 2 unordered
// output -- only ever the last pass of a forward transform (contiguous rows for every plan with >= 2 passes)
template <int LOGL, int C, int R>
cudaError_t launch_tile_lc (int dir, bool jfast, int uio, const TileArgs& a, cudaStream_t stream)
{
    if (uio == 1)
        return (dir > 0 && ! jfast) ? launch_tile_one<LOGL, C, +1, false, 1, R> (a, stream) : cudaErrorInvalidValue;
    if (uio == 2)
        return (dir < 0 && jfast) ? launch_tile_one<LOGL, C, -1, true, 2, R> (a, stream) : cudaErrorInvalidValue;
    if (dir < 0)
        return jfast ? launch_tile_one<LOGL, C, -1, true, 0, R> (a, stream) : launch_tile_one<LOGL, C, -1, false, 0, R> (a, stream);
    return jfast ? launch_tile_one<LOGL, C, +1, true, 0, R> (a, stream) : launch_tile_one<LOGL, C, +1, false, 0, R> (a, stream);
}
template <int LOGL>
cudaError_t launch_tile_l (int C, int dir, bool jfast, int uio, const TileArgs& a, cudaStream_t stream)
{
    // 32 points per thread where the tuning hook asks for it and the geometry exists (512 / 1024 points: 32 x 16, 32 x 32)
    if constexpr (LOGL >= 9)
        if (tile_radix32() != 0)
        {
            if (C == 8)
                return launch_tile_lc<LOGL, 8, 32> (dir, jfast, uio, a, stream);
            if constexpr (LOGL == 9)
                if (C == 16)
                    return launch_tile_lc<LOGL, 16, 32> (dir, jfast, uio, a, stream);
        }
    if (C == 8)
        return launch_tile_lc<LOGL, 8, 16> (dir, jfast, uio, a, stream);
    if constexpr (LOGL <= 9)
        if (C == 16)
            return launch_tile_lc<LOGL, 16, 16> (dir, jfast, uio, a, stream);
    return cudaErrorInvalidValue;
}
} // namespace

cudaError_t launch_tile (int logL, int C, int dir, bool load_j_fast, int uio, const TileArgs& a, cudaStream_t stream)
{
    switch (logL)
    {
        case 6: return launch_tile_l<6> (C, dir, load_j_fast, uio, a, stream);
        case 7: return launch_tile_l<7> (C, dir, load_j_fast, uio, a, stream);
        case 8: return launch_tile_l<8> (C, dir, load_j_fast, uio, a, stream);
        case 9: return launch_tile_l<9> (C, dir, load_j_fast, uio, a, stream);
        case 10: return launch_tile_l<10> (C, dir, load_j_fast, uio, a, stream);
        default: return cudaErrorInvalidValue;
    }
}

// ---- persistent tensor-map TMA tile kernel (tile_tma_kernel): cudaErrorInvalidConfiguration / NotSupported = does not apply ----
namespace
{
int sm_count_cached()
{
    static thread_local int c_dev = -1, c_sms = 148;
    int dev = 0;
    if (cudaGetDevice (&dev) == cudaSuccess && dev != c_dev)
    {
        int n = 0;
        if (cudaDeviceGetAttribute (&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
            c_sms = n;
        c_dev = dev;
    }
    return c_sms;
}
template <int LOGL, int C, int DIR, bool JFAST>
cudaError_t launch_tile_tma_one (const TilePass& p, cudaStream_t stream)
{
    using TT = TileTmaLaunch<LOGL, C, 16>;
    if constexpr (! TT::FITS)
        return cudaErrorInvalidConfiguration;
    else
    {
        TileTmaSide in, out;
        if (! build_tile_tma (p, in, out))
            return cudaErrorInvalidConfiguration;
        TensorMap5 im, om;
        cudaError_t e = tma_make_map5 (in.base, in.dims, in.strides, in.box, im);
        if (e == cudaSuccess)
            e = tma_make_map5 (out.base, out.dims, out.strides, out.box, om);
        if (e != cudaSuccess)
            return e == cudaErrorInvalidValue ? cudaErrorInvalidConfiguration : e; // a shape the encoder rejects: use tile_fft_kernel
        auto kernel = tile_tma_kernel<LOGL, C, DIR, JFAST, 16>;
        static thread_local int attr_dev = -1, resident = 0;
        int dev = 0;
        if ((e = cudaGetDevice (&dev)) != cudaSuccess)
            return e;
        if (dev != attr_dev)
        {
            int per_sm = 0;
            if ((e = cudaFuncSetAttribute (kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TT::SMEM_BYTES)) != cudaSuccess
                || (e = cudaOccupancyMaxActiveBlocksPerMultiprocessor (&per_sm, kernel, TT::THREADS, (size_t) TT::SMEM_BYTES)) != cudaSuccess)
                return e;
            if (per_sm < 1)
                return cudaErrorInvalidConfiguration;
            attr_dev = dev;
            resident = sm_count_cached() * per_sm;
        }
        const long long tiles = (long long) p.args.ntiles * p.args.batch;
        if (tiles <= 0)
            return cudaSuccess;
        kernel<<<(unsigned) (tiles < resident ? tiles : resident), TT::THREADS, TT::SMEM_BYTES, stream>>> (im, om, p.args, in.coords, out.coords);
        count_launch();
        return cudaGetLastError();
    }
}
template <int LOGL, int C>
cudaError_t launch_tile_tma_lc (int dir, const TilePass& p, cudaStream_t stream)
{
    if (dir < 0)
        return p.load_j_fast ? launch_tile_tma_one<LOGL, C, -1, true> (p, stream) : launch_tile_tma_one<LOGL, C, -1, false> (p, stream);
    return p.load_j_fast ? launch_tile_tma_one<LOGL, C, +1, true> (p, stream) : launch_tile_tma_one<LOGL, C, +1, false> (p, stream);
}
} // namespace

cudaError_t launch_tile_tma (int dir, const TilePass& p, cudaStream_t stream)
{
    switch (p.logL * 100 + p.C)
    {
        case 808: return launch_tile_tma_lc<8, 8> (dir, p, stream);
        case 816: return launch_tile_tma_lc<8, 16> (dir, p, stream);
        case 908: return launch_tile_tma_lc<9, 8> (dir, p, stream);
        case 916: return launch_tile_tma_lc<9, 16> (dir, p, stream);
        case 1008: return launch_tile_tma_lc<10, 8> (dir, p, stream);
        default: return cudaErrorInvalidConfiguration;
    }
}

// ---- tile_fft_kernel with the tensor-map L2 prefetch (tile_fft_pf_kernel); cudaErrorInvalidConfiguration = does not apply ----
namespace
{
template <int LOGL, int C, int DIR, bool JFAST>
cudaError_t launch_tile_pf_one (const TilePass& p, int distance, cudaStream_t stream)
{
    using TL = TileLaunch<LOGL, C, 16>;
    TileTmaSide in, out;
    if (! build_tile_tma (p, in, out))
        return cudaErrorInvalidConfiguration;
    TensorMap5 im;
    cudaError_t e = tma_make_map5 (in.base, in.dims, in.strides, in.box, im);
    if (e != cudaSuccess)
        return e == cudaErrorInvalidValue ? cudaErrorInvalidConfiguration : e;
    auto kernel = tile_fft_pf_kernel<LOGL, C, DIR, JFAST, 16>;
    if (TL::SMEM_BYTES > 48 * 1024)
        if ((e = cudaFuncSetAttribute (kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TL::SMEM_BYTES)) != cudaSuccess)
            return e;
    kernel<<<(unsigned) p.args.ntiles * (unsigned) p.args.batch, TL::THREADS, TL::SMEM_BYTES, stream>>> (im, p.args, in.coords, distance);
    count_launch();
    return cudaGetLastError();
}
template <int LOGL, int C>
cudaError_t launch_tile_pf_lc (int dir, const TilePass& p, int distance, cudaStream_t stream)
{
    if (dir < 0)
        return p.load_j_fast ? launch_tile_pf_one<LOGL, C, -1, true> (p, distance, stream) : launch_tile_pf_one<LOGL, C, -1, false> (p, distance, stream);
    return p.load_j_fast ? launch_tile_pf_one<LOGL, C, +1, true> (p, distance, stream) : launch_tile_pf_one<LOGL, C, +1, false> (p, distance, stream);
}
cudaError_t launch_tile_pf (int dir, const TilePass& p, int distance, cudaStream_t stream)
{
    switch (p.logL * 100 + p.C)
    {
        case 708: return launch_tile_pf_lc<7, 8> (dir, p, distance, stream);
        case 716: return launch_tile_pf_lc<7, 16> (dir, p, distance, stream);
        case 808: return launch_tile_pf_lc<8, 8> (dir, p, distance, stream);
        case 816: return launch_tile_pf_lc<8, 16> (dir, p, distance, stream);
        case 908: return launch_tile_pf_lc<9, 8> (dir, p, distance, stream);
        case 916: return launch_tile_pf_lc<9, 16> (dir, p, distance, stream);
        case 1008: return launch_tile_pf_lc<10, 8> (dir, p, distance, stream);
        default: return cudaErrorInvalidConfiguration;
    }
}
} // namespace

// one tile pass: the TMA kernel where the tuning hook allows and the pass can be expressed, else tile_fft_kernel
cudaError_t launch_tile_pass (int dir, const TilePass& p_in, cudaStream_t stream)
{
    TilePass p = p_in;
    if ((tile_stream_mode() & 1) != 0 && p.args.in_policy == POLICY_NORMAL)
        p.args.in_policy = POLICY_STREAM;
    if ((tile_stream_mode() & 2) != 0 && p.args.out_policy == POLICY_NORMAL)
        p.args.out_policy = POLICY_STREAM;
    if (tile_tma_mode() != 0 && tile_radix32() == 0)
    {
        const cudaError_t e = launch_tile_tma (dir, p, stream);
        if (e != cudaErrorInvalidConfiguration && e != cudaErrorNotSupported)
            return e;
        (void) cudaGetLastError();
    }
    if (tile_pf_distance() > 0 && tile_radix32() == 0)
    {
        const cudaError_t e = launch_tile_pf (dir, p, tile_pf_distance(), stream);
        if (e != cudaErrorInvalidConfiguration && e != cudaErrorNotSupported)
            return e;
        (void) cudaGetLastError();
    }
    return launch_tile (p.logL, p.C, dir, p.load_j_fast, p.uio, p.args, stream);
}

cudaError_t launch_dist_barrier (const DistBarrierArgs& a, cudaStream_t stream)
{
    dist_barrier_kernel<0><<<1, 32, 0, stream>>> (a);
    count_launch();
    return cudaGetLastError();
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

} // namespace cfb
