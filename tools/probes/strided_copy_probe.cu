// Memory-pattern ceiling for the strided tile passes of the large transforms (tools/ only, not part of the library):
// copies a [L][S] matrix of float2 tile by tile -- a CTA moves L rows x C adjacent columns (C*8 contiguous bytes per row,
// row stride S*8 bytes) from `in` to the same place in `out`, 16 independent 8-byte accesses per thread, no arithmetic.
// This is exactly the global-memory pattern of tile_fft_kernel's pass A; what it reaches is the most that pass can.
#include <cstdio>
#include <cuda_runtime.h>

template <int C>
__global__ void __launch_bounds__ (1024) tile_copy (const float2* __restrict__ in, float2* __restrict__ out, long long S, int L)
{
    const int tid = threadIdx.x, lt = tid % C, j = tid / C, T = blockDim.x / C;
    const long long base = (long long) blockIdx.x * C + lt;
    float2 v[16];
#pragma unroll
    for (int m = 0; m < 16; ++m)
        v[m] = __ldcs (in + base + (long long) (j + m * T) * S);
#pragma unroll
    for (int m = 0; m < 16; ++m)
        out[base + (long long) (j + m * T) * S] = v[m];
}
// pattern of the contiguous-row pass: a CTA reads C rows of L contiguous values and writes them transposed, C adjacent values
// (C*8 bytes) per output row of stride S*8 bytes
template <int C>
__global__ void __launch_bounds__ (1024) row_to_tile_copy (const float2* __restrict__ in, float2* __restrict__ out, long long S, int L)
{
    extern __shared__ float2 sm[];
    const int tid = threadIdx.x, T = blockDim.x / C;
    const int ltA = tid / T, jA = tid % T, ltB = tid % C, jB = tid / C;
    const long long row0 = (long long) blockIdx.x * C;
#pragma unroll
    for (int m = 0; m < 16; ++m)
        sm[ltA * (L + 1) + jA + m * T] = __ldcs (in + (row0 + ltA) * L + jA + m * T);
    __syncthreads();
#pragma unroll
    for (int m = 0; m < 16; ++m)
        out[row0 + ltB + (long long) (jB + m * T) * S] = sm[ltB * (L + 1) + jB + m * T];
}
__global__ void linear_copy (const float4* __restrict__ in, float4* __restrict__ out, long long n)
{
    for (long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x; i < n; i += (long long) gridDim.x * blockDim.x)
        out[i] = in[i];
}

template <int C>
void run (const float2* in, float2* out, long long N, int L)
{
    const long long S = N / L;
    const int threads = (L / 16) * C;
    cudaEvent_t e0, e1;
    cudaEventCreate (&e0);
    cudaEventCreate (&e1);
    for (int rep = 0; rep < 2; ++rep)
    {
        cudaEventRecord (e0);
        for (int i = 0; i < 5; ++i)
            tile_copy<C><<<(unsigned) (S / C), threads>>> (in, out, S, L);
        cudaEventRecord (e1);
        cudaEventSynchronize (e1);
    }
    float ms = 0;
    cudaEventElapsedTime (&ms, e0, e1);
    ms /= 5;
    printf ("tile copy  L=%4d  C=%2d (%3d-byte pieces, %4d threads): %.3f ms  %.0f GB/s (read+write)\n", L, C, C * 8, threads, ms, 2.0 * N * 8 / ms / 1e6);
}

template <int C>
void run_rows (const float2* in, float2* out, long long N, int L)
{
    const long long S = N / L; // rows
    const int threads = (L / 16) * C;
    const size_t smem = (size_t) C * (L + 1) * 8;
    cudaFuncSetAttribute (row_to_tile_copy<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    cudaEvent_t e0, e1;
    cudaEventCreate (&e0);
    cudaEventCreate (&e1);
    for (int rep = 0; rep < 2; ++rep)
    {
        cudaEventRecord (e0);
        for (int i = 0; i < 5; ++i)
            row_to_tile_copy<C><<<(unsigned) (S / C), threads, smem>>> (in, out, S, L);
        cudaEventRecord (e1);
        cudaEventSynchronize (e1);
    }
    float ms = 0;
    cudaEventElapsedTime (&ms, e0, e1);
    ms /= 5;
    printf ("rows -> tile  L=%4d  C=%2d (linear %d-byte rows in, %3d-byte pieces out, %4d threads, %3zu KB smem): %.3f ms  %.0f GB/s (read+write)\n", L, C, L * 8, C * 8, threads,
            smem / 1024, ms, 2.0 * N * 8 / ms / 1e6);
}

int main()
{
    const long long N = 1LL << 28;
    float2 *in, *out;
    cudaMalloc (&in, N * 8);
    cudaMalloc (&out, N * 8);
    cudaMemset (in, 0, N * 8);
    cudaEvent_t e0, e1;
    cudaEventCreate (&e0);
    cudaEventCreate (&e1);
    for (int rep = 0; rep < 2; ++rep)
    {
        cudaEventRecord (e0);
        for (int i = 0; i < 5; ++i)
            linear_copy<<<148 * 16, 512>>> ((const float4*) in, (float4*) out, N / 2);
        cudaEventRecord (e1);
        cudaEventSynchronize (e1);
    }
    float ms = 0;
    cudaEventElapsedTime (&ms, e0, e1);
    printf ("linear copy: %.3f ms  %.0f GB/s (read+write)\n", ms / 5, 2.0 * N * 8 / (ms / 5) / 1e6);
    for (int L : { 512, 1024 })
    {
        run<4> (in, out, N, L);
        run<8> (in, out, N, L);
        run<16> (in, out, N, L);
        if (L == 512)
            run<32> (in, out, N, L);
    }
    run_rows<8> (in, out, N, 1024);
    run_rows<16> (in, out, N, 1024);
    run_rows<8> (in, out, N, 512);
    run_rows<16> (in, out, N, 512);
    run_rows<32> (in, out, N, 512);
    run_rows<16> (in, out, N, 256);
    run_rows<32> (in, out, N, 256);
    printf ("%s\n", cudaGetErrorString (cudaDeviceSynchronize()));
    return 0;
}
