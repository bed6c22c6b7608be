// Instantiations and launcher of the one-pass cluster transform (cluster_kernels.cuh): tensor-map creation through the
// driver entry point (no libcuda link dependency), cluster launch attributes, co-resident cluster count.
#include <cuda.h>

#include <cstdint>

#include "cluster_kernels.cuh"
#include "dispatch.h"

namespace cfb
{
namespace
{
using EncodeTiledFn = CUresult (*) (CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                    CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled()
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

// 4-D view of `batch` transforms of N complex values: [batch][c : 512][g : G][32 floats = one 128-byte line]
template <int LOGG>
cudaError_t make_input_map (const float* in, long long in_stride, int batch, TensorMap4& out)
{
    using CG = ClusterGeo<LOGG>;
    static_assert (sizeof (TensorMap4) == sizeof (CUtensorMap) && alignof (TensorMap4) >= alignof (CUtensorMap), "TensorMap4 must mirror CUtensorMap");
    const EncodeTiledFn enc = encode_tiled();
    if (enc == nullptr)
        return cudaErrorNotSupported;
    const cuuint64_t dims[4] = { 32, (cuuint64_t) CG::G, (cuuint64_t) CG::LC, (cuuint64_t) batch };
    const cuuint64_t strides[3] = { 128, 128ull * CG::G, (cuuint64_t) in_stride * 4ull };
    const cuuint32_t box[4] = { 32, 1, (cuuint32_t) CG::TMA_ROWS, 1 };
    const cuuint32_t estr[4] = { 1, 1, 1, 1 };
    const CUresult r = enc (reinterpret_cast<CUtensorMap*> (&out), CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*> (in), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

// 3-D view of the natural-order output: [batch][N / 512 rows][1024 floats]; one CTA's part of a spectrum is the box
// { 2 RUN floats, 16 G rows } at column 2 RUN g: bins RUN g + r + 512 (kb + G ka)
template <int LOGG>
cudaError_t make_output_map (float* out, long long out_stride, int batch, TensorMap4& map)
{
    using CG = ClusterGeo<LOGG>;
    const EncodeTiledFn enc = encode_tiled();
    if (enc == nullptr)
        return cudaErrorNotSupported;
    const cuuint64_t dims[3] = { 1024, (cuuint64_t) (16 * CG::G), (cuuint64_t) batch };
    const cuuint64_t strides[2] = { 4096, (cuuint64_t) out_stride * 4ull };
    const cuuint32_t box[3] = { (cuuint32_t) (2 * CG::RUN), (cuuint32_t) (16 * CG::G), 1 };
    const cuuint32_t estr[3] = { 1, 1, 1 };
    const CUresult r = enc (reinterpret_cast<CUtensorMap*> (&map), CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, out, dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

template <int LOGG, int DIR, int LOGW>
cudaError_t launch_cluster_one (const float* in, long long in_stride, const ClusterArgs& a, cudaStream_t stream)
{
    using CG = ClusterGeo<LOGG>;
    auto kernel = cluster_fft_kernel<LOGG, DIR, LOGW>;
    static thread_local int c_dev = -1, c_clusters = 0;
    int dev = 0;
    cudaError_t e = cudaGetDevice (&dev);
    if (e != cudaSuccess)
        return e;
    cudaLaunchConfig_t cfg {};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CG::G;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.blockDim = dim3 (CG::THREADS);
    cfg.dynamicSmemBytes = CG::SMEM_BYTES;
    cfg.stream = stream;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (dev != c_dev)
    {
        if ((e = cudaFuncSetAttribute (kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CG::SMEM_BYTES)) != cudaSuccess)
            return e;
        if (CG::G > 8 && (e = cudaFuncSetAttribute (kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1)) != cudaSuccess)
            return e;
        int n = 0;
        cfg.gridDim = dim3 (CG::G * 64);
        if ((e = cudaOccupancyMaxActiveClusters (&n, kernel, &cfg)) != cudaSuccess)
            return e;
        if (n < 1)
            return cudaErrorInvalidConfiguration; // this cluster size cannot be co-scheduled here: the caller uses the multi-pass path
        c_dev = dev;
        c_clusters = n;
    }
    if (a.batch <= 0)
        return cudaSuccess;
    TensorMap4 tm, om;
    if ((e = make_input_map<LOGG> (in, in_stride, a.batch, tm)) != cudaSuccess)
        return e;
    if (LOGW != 0)
        om = tm; // unused by the kernel
    else if ((e = make_output_map<LOGG> (a.out, a.out_stride, a.batch, om)) != cudaSuccess)
        return e;
    const int clusters = a.batch < c_clusters ? a.batch : c_clusters;
    cfg.gridDim = dim3 ((unsigned) (clusters * CG::G));
    e = cudaLaunchKernelEx (&cfg, kernel, tm, om, a);
    count_launch();
    return e != cudaSuccess ? e : cudaGetLastError();
}

template <int LOGG>
cudaError_t launch_cluster_g (int dir, int logW, const float* in, long long in_stride, const ClusterArgs& a, cudaStream_t stream)
{
    if (dir > 0)
        return logW == 0 ? launch_cluster_one<LOGG, +1, 0> (in, in_stride, a, stream) : cudaErrorInvalidConfiguration;
    if (logW == 3)
        return launch_cluster_one<LOGG, -1, 3> (in, in_stride, a, stream);
    return logW == 0 ? launch_cluster_one<LOGG, -1, 0> (in, in_stride, a, stream) : cudaErrorInvalidConfiguration;
}
} // namespace

bool has_cluster (int logN) { return logN >= 15 && logN <= 17; }

// cudaErrorInvalidConfiguration / cudaErrorNotSupported = this case does not apply here (the caller falls back to the tile passes)
cudaError_t launch_cluster_fft (int logN, int dir, int logW, const float* in, long long in_stride, const ClusterArgs& a, cudaStream_t stream)
{
    switch (logN)
    {
        case 15: return launch_cluster_g<2> (dir, logW, in, in_stride, a, stream);
        case 16: return launch_cluster_g<3> (dir, logW, in, in_stride, a, stream);
        case 17: return launch_cluster_g<4> (dir, logW, in, in_stride, a, stream);
        default: return cudaErrorInvalidConfiguration;
    }
}
} // namespace cfb
