// L2 probe for the chunked large-transform schedules (tools/ only, not part of the product):
//   1. read bandwidth of a buffer of W MiB read repeatedly by all SMs (L2-hit bandwidth while W fits, HBM beyond)
//   2. producer -> consumer through a ring slot: kernel P streams `total` bytes from HBM and writes slot (W MiB), kernel C
//      reads the slot and streams to HBM -- the pattern of pass B -> pass C; time per byte vs a plain copy
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/probes/l2_probe tools/probes/l2_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf ("%s: %s\n", #x, cudaGetErrorString (e)); return 1; } } while (0)

__global__ void read_kernel (const float4* __restrict__ p, size_t n4, int reps, float* sink)
{
    float acc = 0.f;
    for (int r = 0; r < reps; ++r)
        for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t) gridDim.x * blockDim.x)
        {
            float4 v;
            asm volatile ("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p + i));
            acc += v.x + v.y + v.z + v.w;
        }
    if (acc == 123.456f)
        *sink = acc;
}
__global__ void copy_kernel (const float4* __restrict__ in, float4* __restrict__ out, size_t n4)
{
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t) gridDim.x * blockDim.x)
        out[i] = in[i];
}

int main()
{
    cudaDeviceProp prop;
    CK (cudaGetDeviceProperties (&prop, 0));
    printf ("device %s  SMs %d  l2CacheSize %.1f MiB  persistingL2CacheMaxSize %.1f MiB  accessPolicyMaxWindowSize %.1f MiB\n", prop.name, prop.multiProcessorCount,
            prop.l2CacheSize / 1048576.0, prop.persistingL2CacheMaxSize / 1048576.0, prop.accessPolicyMaxWindowSize / 1048576.0);
    const size_t big = (size_t) 2 << 30;
    float4 *a, *b, *ring;
    float* sink;
    CK (cudaMalloc (&a, big));
    CK (cudaMalloc (&b, big));
    CK (cudaMalloc (&ring, (size_t) 512 << 20));
    CK (cudaMalloc (&sink, 4));
    CK (cudaMemset (a, 0, big));
    CK (cudaMemset (b, 0, big));
    CK (cudaMemset (ring, 0, (size_t) 512 << 20));
    cudaEvent_t e0, e1;
    CK (cudaEventCreate (&e0));
    CK (cudaEventCreate (&e1));
    const int grid = prop.multiProcessorCount * 8, block = 256;
    float ms;
    printf ("== 1. repeated reads of W MiB (all SMs)\n");
    for (int w : { 4, 8, 16, 24, 32, 48, 64, 80, 96, 112, 128, 192, 256, 512 })
    {
        const size_t n4 = ((size_t) w << 20) / 16;
        const int reps = 2048 / w < 4 ? 4 : 2048 / w;
        read_kernel<<<grid, block>>> (ring, n4, 2, sink);
        CK (cudaEventRecord (e0));
        read_kernel<<<grid, block>>> (ring, n4, reps, sink);
        CK (cudaEventRecord (e1));
        CK (cudaEventSynchronize (e1));
        CK (cudaEventElapsedTime (&ms, e0, e1));
        printf ("   W = %4d MiB  %8.1f GB/s\n", w, (double) n4 * 16 * reps / ms / 1e6);
    }
    printf ("== 2. plain copy 2 GiB -> 2 GiB\n");
    for (int it = 0; it < 3; ++it)
    {
        CK (cudaEventRecord (e0));
        copy_kernel<<<grid, block>>> (a, b, big / 16);
        CK (cudaEventRecord (e1));
        CK (cudaEventSynchronize (e1));
        CK (cudaEventElapsedTime (&ms, e0, e1));
        printf ("   %.3f ms  %.1f GB/s (read+write)\n", ms, 2.0 * big / ms / 1e6);
    }
    printf ("== 3. HBM -> ring slot -> HBM in chunks of W MiB over L lanes (two streams / slots alternate); total 2 GiB each way\n");
    cudaStream_t st[4];
    for (auto& s : st)
        CK (cudaStreamCreateWithFlags (&s, cudaStreamNonBlocking));
    for (int lanes : { 1, 2, 3 })
        for (int w : { 4, 8, 16, 32, 64 })
        {
            const size_t cb = (size_t) w << 20, n4 = cb / 16;
            const int chunks = (int) (big / cb);
            for (int it = 0; it < 2; ++it)
            {
                CK (cudaDeviceSynchronize());
                CK (cudaEventRecord (e0, st[0]));
                for (int l = 1; l < lanes; ++l)
                    CK (cudaStreamWaitEvent (st[l], e0, 0));
                for (int c = 0; c < chunks; ++c)
                {
                    const int l = c % lanes;
                    float4* slot = ring + (size_t) l * n4;
                    copy_kernel<<<grid, block, 0, st[l]>>> (a + (size_t) c * n4, slot, n4);
                    copy_kernel<<<grid, block, 0, st[l]>>> (slot, b + (size_t) c * n4, n4);
                }
                cudaEvent_t j;
                for (int l = 1; l < lanes; ++l)
                {
                    CK (cudaEventCreateWithFlags (&j, cudaEventDisableTiming));
                    CK (cudaEventRecord (j, st[l]));
                    CK (cudaStreamWaitEvent (st[0], j, 0));
                    CK (cudaEventDestroy (j));
                }
                CK (cudaEventRecord (e1, st[0]));
                CK (cudaEventSynchronize (e1));
                CK (cudaEventElapsedTime (&ms, e0, e1));
            }
            printf ("   lanes %d  W = %3d MiB  %d chunks  %.3f ms  = %.1f GB/s of HBM-facing traffic (2 x 2 GiB)\n", lanes, w, chunks, ms, 2.0 * big / ms / 1e6);
        }
    printf ("done\n");
    return 0;
}
