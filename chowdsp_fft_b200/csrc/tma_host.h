// Host side of tma.cuh: tensor-map creation through the driver entry point (no libcuda link dependency).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include "tma.cuh"

namespace cfb
{
using EncodeTiledFn = CUresult (*) (CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                    CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline EncodeTiledFn tma_encode_tiled()
{
    static EncodeTiledFn fn = []() -> EncodeTiledFn
    {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
        if (cudaGetDriverEntryPoint ("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
        {
            (void) cudaGetLastError();
            return nullptr;
        }
        return reinterpret_cast<EncodeTiledFn> (p);
    }();
    return fn;
}

// fp32 tensor of rank 5: dims in floats (innermost first), strides in bytes for dimensions 1..4, box in elements
inline cudaError_t tma_make_map5 (const void* base, const unsigned long long (&dims)[5], const unsigned long long (&strides)[4], const unsigned (&box)[5], TensorMap5& out)
{
    static_assert (sizeof (TensorMap5) == sizeof (CUtensorMap) && alignof (TensorMap5) >= alignof (CUtensorMap), "TensorMap5 must mirror CUtensorMap");
    const EncodeTiledFn enc = tma_encode_tiled();
    if (enc == nullptr)
        return cudaErrorNotSupported;
    cuuint64_t d[5], st[4];
    cuuint32_t b[5], es[5] = { 1, 1, 1, 1, 1 };
    for (int i = 0; i < 5; ++i)
    {
        d[i] = dims[i];
        b[i] = box[i];
    }
    for (int i = 0; i < 4; ++i)
        st[i] = strides[i];
    const CUresult r = enc (reinterpret_cast<CUtensorMap*> (&out), CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<void*> (base), d, st, b, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}
} // namespace cfb
