// Instantiates the persistent TMA-pipelined transform (pipe_kernels.cuh) for ONE complex length 2^CFB_LOGM.
#ifndef CFB_LOGM
#error "compile with -DCFB_LOGM=<13|14>"
#endif
#include "dispatch.h"
#include "pipe_kernels.cuh"

namespace cfb
{
namespace
{
int sm_count()
{
    static thread_local int cached_dev = -1, cached_sms = 0;
    int dev = 0;
    if (cudaGetDevice (&dev) != cudaSuccess)
        return 148;
    if (dev != cached_dev)
    {
        int n = 0;
        if (cudaDeviceGetAttribute (&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
            n = 148;
        cached_dev = dev;
        cached_sms = n;
    }
    return cached_sms;
}

template <int KIND, int LOGW>
cudaError_t launch_pipe_one (const FftArgs& a, cudaStream_t stream)
{
    using P = PipeGeo<CFB_LOGM, KIND, LOGW>;
    auto kernel = pipe_kernel<CFB_LOGM, KIND, LOGW>;
    const cudaError_t e = cudaFuncSetAttribute (kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, P::SMEM_BYTES);
    if (e != cudaSuccess)
        return e;
    if (a.batch <= 0)
        return cudaSuccess;
    const int ctas = sm_count() * P::CTAS_PER_SM;
    const unsigned grid = (unsigned) (a.batch < ctas ? a.batch : ctas);
    kernel<<<grid, P::T, P::SMEM_BYTES, stream>>> (a);
    count_launch();
    return cudaGetLastError();
}
} // namespace

#define CFB_CAT2(a, b) a##b
#define CFB_CAT(a, b) CFB_CAT2 (a, b)

// logW: 0 = ordered, 3 = 8-lane unordered layout; `a` must describe a plain batch with 16-byte aligned input rows
cudaError_t CFB_CAT (launch_pipe_, CFB_LOGM) (int kind, int logW, const FftArgs& a, cudaStream_t stream)
{
    switch (kind * 4 + logW)
    {
        case C2C_FWD * 4: return launch_pipe_one<C2C_FWD, 0> (a, stream);
        case C2C_BWD * 4: return launch_pipe_one<C2C_BWD, 0> (a, stream);
        case R2C * 4: return launch_pipe_one<R2C, 0> (a, stream);
        case C2R * 4: return launch_pipe_one<C2R, 0> (a, stream);
        case C2C_FWD * 4 + 3: return launch_pipe_one<C2C_FWD, 3> (a, stream);
        case C2C_BWD * 4 + 3: return launch_pipe_one<C2C_BWD, 3> (a, stream);
        case R2C * 4 + 3: return launch_pipe_one<R2C, 3> (a, stream);
        case C2R * 4 + 3: return launch_pipe_one<C2R, 3> (a, stream);
        default: return cudaErrorInvalidValue;
    }
}
} // namespace cfb
