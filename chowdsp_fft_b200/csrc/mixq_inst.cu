// Instantiations and launcher of the Q x 2^p mixed-radix kernels (mixq_kernels.cuh).
#include "dispatch.h"
#include "mixq_kernels.cuh"

namespace cfb
{
namespace
{
template <int LOGP, int Q>
cudaError_t launch_mixq_one (const MixQArgs& a, cudaStream_t stream)
{
    using X = MixQGeo<LOGP, Q>;
    if constexpr (X::M > kMixedMaxM || X::THREADS > 1024 || X::SMEM_BYTES > 227 * 1024)
        return cudaErrorInvalidConfiguration;
    else
    {
        const bool fast = (a.kind == C2C_FWD || a.kind == C2C_BWD) && a.W == 0;
        auto kernel = fast ? mixq_kernel<LOGP, Q, 0> : mixq_kernel<LOGP, Q, 1>;
        if (X::SMEM_BYTES > 48 * 1024)
        {
            const cudaError_t e = cudaFuncSetAttribute (kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, X::SMEM_BYTES);
            if (e != cudaSuccess)
                return e;
        }
        if (a.batch <= 0)
            return cudaSuccess;
        const long long ctas = ((long long) a.batch + X::SLOTS - 1) / X::SLOTS;
        kernel<<<(unsigned) ctas, X::THREADS, X::SMEM_BYTES, stream>>> (a);
        count_launch();
        return cudaGetLastError();
    }
}
template <int LOGP>
cudaError_t launch_mixq_p (int Q, const MixQArgs& a, cudaStream_t stream)
{
    switch (Q)
    {
        case 3: return launch_mixq_one<LOGP, 3> (a, stream);
        case 5: return launch_mixq_one<LOGP, 5> (a, stream);
        case 9: return launch_mixq_one<LOGP, 9> (a, stream);
        case 15: return launch_mixq_one<LOGP, 15> (a, stream);
        default: return cudaErrorInvalidConfiguration;
    }
}
} // namespace

// cudaErrorInvalidConfiguration: no such instance (the caller falls back to the generic kernel)
cudaError_t launch_mixq (int logP, int Q, const MixQArgs& a, cudaStream_t stream)
{
    switch (logP)
    {
        case 4: return launch_mixq_p<4> (Q, a, stream);
        case 5: return launch_mixq_p<5> (Q, a, stream);
        case 6: return launch_mixq_p<6> (Q, a, stream);
        case 7: return launch_mixq_p<7> (Q, a, stream);
        case 8: return launch_mixq_p<8> (Q, a, stream);
        case 9: return launch_mixq_p<9> (Q, a, stream);
        case 10: return launch_mixq_p<10> (Q, a, stream);
        case 11: return launch_mixq_p<11> (Q, a, stream);
        case 12: return launch_mixq_p<12> (Q, a, stream);
        default: return cudaErrorInvalidConfiguration;
    }
}
} // namespace cfb
