// Instantiates the single-kernel transform for ONE complex length 2^CFB_LOGM (all kinds, ordered and
// unordered).  Compiled once per size so the sizes build in parallel.
#ifndef CFB_LOGM
#error "compile with -DCFB_LOGM=<4..14>"
#endif
#include "dispatch.h"

namespace cfb
{
namespace
{
template <int KIND, bool UNORD>
cudaError_t launch_one (const FftArgs& a, cudaStream_t stream)
{
    using L = Launch<CFB_LOGM, kRadix>;
    auto kernel = fft_kernel<CFB_LOGM, kRadix, KIND, UNORD>;
    constexpr int smem_bytes = UNORD ? L::SMEM_BYTES_UNORD : L::SMEM_BYTES;
    if (smem_bytes > 48 * 1024)
    {
        const cudaError_t e = cudaFuncSetAttribute (kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
        if (e != cudaSuccess)
            return e;
    }
    if (a.batch <= 0)
        return cudaSuccess;
    const unsigned grid = (unsigned) (((long long) a.batch + L::PER_CTA - 1) / L::PER_CTA);
    kernel<<<grid, L::THREADS, smem_bytes, stream>>> (a);
    count_launch();
    return cudaGetLastError();
}
} // namespace

#define CFB_CAT2(a, b) a##b
#define CFB_CAT(a, b) CFB_CAT2 (a, b)

cudaError_t CFB_CAT (launch_fft_, CFB_LOGM) (int kind, bool unordered, const FftArgs& a, cudaStream_t stream)
{
    switch (kind * 2 + (unordered ? 1 : 0))
    {
        case 0: return launch_one<C2C_FWD, false> (a, stream);
        case 1: return launch_one<C2C_FWD, true> (a, stream);
        case 2: return launch_one<C2C_BWD, false> (a, stream);
        case 3: return launch_one<C2C_BWD, true> (a, stream);
        case 4: return launch_one<R2C, false> (a, stream);
        case 5: return launch_one<R2C, true> (a, stream);
        case 6: return launch_one<C2R, false> (a, stream);
        case 7: return launch_one<C2R, true> (a, stream);
        default: return cudaErrorInvalidValue;
    }
}

int CFB_CAT (stage_twiddle_len_, CFB_LOGM)() { return Geo<CFB_LOGM, kRadix>::TW_LEN; }
void CFB_CAT (fill_stage_twiddles_, CFB_LOGM) (float2* tw) { fill_stage_twiddles<CFB_LOGM, kRadix> (tw); }
} // namespace cfb
