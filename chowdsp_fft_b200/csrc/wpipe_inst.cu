// Instantiates the warp-pipelined transform (wpipe_kernel, pipe_kernels.cuh) for ONE complex length 2^CFB_LOGM whose
// transforms are owned by a single warp: 2^10 points at 32 per thread, 2^9 at 16.
#ifndef CFB_LOGM
#error "compile with -DCFB_LOGM=<9|10>"
#endif
#include "dispatch.h"
#include "pipe_kernels.cuh"

namespace cfb
{
namespace
{
constexpr int kR = CFB_LOGM == 10 ? 32 : 16;

template <int KIND, int LOGW>
cudaError_t launch_wpipe_one (int warps, const FftArgs& a, cudaStream_t stream)
{
    using WP = WPipeGeo<CFB_LOGM, kR, LOGW>;
    auto kernel = wpipe_kernel<CFB_LOGM, kR, KIND, LOGW>;
    if (warps <= 0 || warps > WP::MAX_WARPS)
        warps = WP::MAX_WARPS;
    const int smem_bytes = WP::smem_bytes (warps);
    if (a.batch <= 0)
        return cudaSuccess;
    // resident CTAs per SM for this CTA shape; the attribute and the occupancy query are cached per thread
    static thread_local int c_dev = -1, c_warps = -1, c_resident = 0;
    int dev = 0;
    cudaError_t e = cudaGetDevice (&dev);
    if (e != cudaSuccess)
        return e;
    if (dev != c_dev || warps != c_warps)
    {
        int sms = 0, per_sm = 0;
        if ((e = cudaFuncSetAttribute (kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WP::smem_bytes (WP::MAX_WARPS))) != cudaSuccess
            || (e = cudaDeviceGetAttribute (&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess
            || (e = cudaOccupancyMaxActiveBlocksPerMultiprocessor (&per_sm, kernel, warps * 32, (size_t) smem_bytes)) != cudaSuccess)
            return e;
        if (per_sm < 1)
            return cudaErrorInvalidConfiguration;
        c_dev = dev;
        c_warps = warps;
        c_resident = sms * per_sm;
    }
    const long long ctas = ((long long) a.batch + warps - 1) / warps;
    kernel<<<(unsigned) (ctas < c_resident ? ctas : c_resident), warps * 32, smem_bytes, stream>>> (a);
    count_launch();
    return cudaGetLastError();
}
} // namespace

namespace
{
template <int HQ>
cudaError_t launch_wistft_one (int warps, const FftArgs& a, cudaStream_t stream)
{
    using WP = WPipeGeo<CFB_LOGM, kR, 0>;
    using WI = WIstftGeo<CFB_LOGM, kR>;
    auto kernel = wistft_kernel<CFB_LOGM, kR, HQ>;
    if (warps <= 0 || warps > WI::MAX_WARPS)
        warps = WI::MAX_WARPS;
    const int smem_bytes = WP::smem_bytes (warps);
    const long long items = (long long) (a.batch / a.inner) * a.nseg;
    if (items <= 0)
        return cudaSuccess;
    static thread_local int c_dev = -1, c_warps = -1, c_resident = 0;
    int dev = 0;
    cudaError_t e = cudaGetDevice (&dev);
    if (e != cudaSuccess)
        return e;
    if (dev != c_dev || warps != c_warps)
    {
        int sms = 0, per_sm = 0;
        if ((e = cudaFuncSetAttribute (kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WP::smem_bytes (WI::MAX_WARPS))) != cudaSuccess
            || (e = cudaDeviceGetAttribute (&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess
            || (e = cudaOccupancyMaxActiveBlocksPerMultiprocessor (&per_sm, kernel, warps * 32, (size_t) smem_bytes)) != cudaSuccess)
            return e;
        if (per_sm < 1)
            return cudaErrorInvalidConfiguration;
        c_dev = dev;
        c_warps = warps;
        c_resident = sms * per_sm;
    }
    const long long ctas = (items + warps - 1) / warps;
    kernel<<<(unsigned) (ctas < c_resident ? ctas : c_resident), warps * 32, smem_bytes, stream>>> (a);
    count_launch();
    return cudaGetLastError();
}
} // namespace

#define CFB_CAT2(a, b) a##b
#define CFB_CAT(a, b) CFB_CAT2 (a, b)

// kind: R2C or C2C_FWD; logW: 0 = ordered, 2 / 3 = unordered output layouts; warps per CTA (<= 0: as many as fit one
// SM); every transform's input row must be 16-byte aligned
cudaError_t CFB_CAT (launch_wpipe_, CFB_LOGM) (int kind, int logW, int warps, const FftArgs& a, cudaStream_t stream)
{
    switch (kind * 4 + logW)
    {
        case C2C_FWD * 4: return launch_wpipe_one<C2C_FWD, 0> (warps, a, stream);
        case C2C_FWD * 4 + 2: return launch_wpipe_one<C2C_FWD, 2> (warps, a, stream);
        case C2C_FWD * 4 + 3: return launch_wpipe_one<C2C_FWD, 3> (warps, a, stream);
        case R2C * 4: return launch_wpipe_one<R2C, 0> (warps, a, stream);
        case R2C * 4 + 2: return launch_wpipe_one<R2C, 2> (warps, a, stream);
        case R2C * 4 + 3: return launch_wpipe_one<R2C, 3> (warps, a, stream);
        default: return cudaErrorInvalidValue;
    }
}

// warp-pipelined overlap-add synthesis (wistft_kernel): hop = 64 hq floats with hq = R/2, R/4 or R/8 (hop = N/2, N/4, N/8);
// args as launch_istft (ordered spectra only), args.seg_frames / args.nseg = segmentation
cudaError_t CFB_CAT (launch_wistft_, CFB_LOGM) (int hq, int warps, const FftArgs& a, cudaStream_t stream)
{
    if (hq == kR / 2)
        return launch_wistft_one<kR / 2> (warps, a, stream);
    if (hq == kR / 4)
        return launch_wistft_one<kR / 4> (warps, a, stream);
    if (hq == kR / 8)
        return launch_wistft_one<kR / 8> (warps, a, stream);
    return cudaErrorInvalidValue;
}
int CFB_CAT (wistft_warps_, CFB_LOGM)() { return WIstftGeo<CFB_LOGM, kR>::MAX_WARPS; }
} // namespace cfb
