// PCIe ceilings for SM-issued traffic (tools/ only, not part of the library): a kernel that reads and / or writes page-locked,
// device-mapped HOST memory directly, against the copy engines on the same buffers.  Question behind it: bench.py's e2e line
// stages host batches through device memory (H2D copy || kernel || D2H copy) and sits on the copy engines' duplex ceiling
// (about 47 GB/s each way); would a transform kernel that loads its input from and stores its result to host memory itself move
// more?  Variants: default / write-combined host allocation, 16-byte accesses, grid sized like the batched FFT (592 CTAs x 256).
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_read (const float4* __restrict__ in, float4* sink, size_t n)
{
    float4 acc = make_float4 (0, 0, 0, 0);
    for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x)
    {
        const float4 v = __ldcs (in + i);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    if (acc.x == 12345.678f)
        sink[0] = acc;
}
__global__ void k_write (float4* __restrict__ out, size_t n)
{
    for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x)
        __stcs (out + i, make_float4 (1.f, 2.f, 3.f, (float) i));
}
// each thread keeps 8 x 16 bytes in flight (like a transform kernel that loads its whole input before it stores)
__global__ void k_copy (const float4* __restrict__ in, float4* __restrict__ out, size_t n)
{
    const size_t stride = (size_t) gridDim.x * blockDim.x;
    for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += 8 * stride)
    {
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u)
            if (i + u * stride < n)
                v[u] = __ldcs (in + i + u * stride);
#pragma unroll
        for (int u = 0; u < 8; ++u)
            if (i + u * stride < n)
                __stcs (out + i + u * stride, v[u]);
    }
}

static float timed (cudaStream_t s, void (*f) (cudaStream_t, void*), void* ctx)
{
    cudaEvent_t e0, e1;
    cudaEventCreate (&e0);
    cudaEventCreate (&e1);
    f (s, ctx); // warm-up
    cudaStreamSynchronize (s);
    float best = 1e30f;
    for (int r = 0; r < 3; ++r)
    {
        cudaEventRecord (e0, s);
        f (s, ctx);
        cudaEventRecord (e1, s);
        cudaEventSynchronize (e1);
        float ms;
        cudaEventElapsedTime (&ms, e0, e1);
        best = ms < best ? ms : best;
    }
    return best;
}

struct Ctx
{
    float4 *hin, *hout, *din, *dout, *dhin, *dhout;
    size_t n;
    int grid;
    cudaStream_t s2;
    cudaEvent_t ev, ev2;
};

int main()
{
    const size_t bytes = 1ull << 30, n = bytes / 16;
    for (int wc = 0; wc < 2; ++wc)
    {
        Ctx c {};
        c.n = n;
        cudaHostAlloc ((void**) &c.hin, bytes, cudaHostAllocMapped | (wc ? cudaHostAllocWriteCombined : 0));
        cudaHostAlloc ((void**) &c.hout, bytes, cudaHostAllocMapped);
        cudaHostGetDevicePointer ((void**) &c.dhin, c.hin, 0);
        cudaHostGetDevicePointer ((void**) &c.dhout, c.hout, 0);
        cudaMalloc ((void**) &c.din, bytes);
        cudaMalloc ((void**) &c.dout, bytes);
        for (size_t i = 0; i < n; i += 4096 / 16)
            c.hin[i] = make_float4 (1, 2, 3, 4);
        cudaStream_t s;
        cudaStreamCreate (&s);
        cudaStreamCreate (&c.s2);
        cudaEventCreate (&c.ev);
        cudaEventCreate (&c.ev2);
        printf ("== host input allocation: %s, 1 GiB each way\n", wc ? "write-combined" : "default (cached)");
        for (int grid : { 148, 592, 2368 })
        {
            c.grid = grid;
            const float r = timed (s, [] (cudaStream_t st, void* p) { Ctx* c = (Ctx*) p; k_read<<<c->grid, 256, 0, st>>> (c->dhin, c->dout, c->n); }, &c);
            const float w = timed (s, [] (cudaStream_t st, void* p) { Ctx* c = (Ctx*) p; k_write<<<c->grid, 256, 0, st>>> (c->dhout, c->n); }, &c);
            const float cp = timed (s, [] (cudaStream_t st, void* p) { Ctx* c = (Ctx*) p; k_copy<<<c->grid, 256, 0, st>>> (c->dhin, c->dhout, c->n); }, &c);
            // kernel reads host memory while a copy engine drains device -> host
            const float mix = timed (s, [] (cudaStream_t st, void* p) {
                Ctx* c = (Ctx*) p;
                cudaEventRecord (c->ev, st);
                cudaStreamWaitEvent (c->s2, c->ev, 0);
                cudaMemcpyAsync (c->hout, c->dout, c->n * 16, cudaMemcpyDeviceToHost, c->s2);
                k_read<<<c->grid, 256, 0, st>>> (c->dhin, c->dout, c->n);
                cudaEventRecord (c->ev2, c->s2);
                cudaStreamWaitEvent (st, c->ev2, 0); }, &c);
            printf ("grid %4d x 256: SM read %6.1f GB/s | SM write %6.1f GB/s | SM copy host->host %6.1f GB/s each way | SM read + CE D2H %6.1f GB/s each way\n",
                    grid, bytes / r / 1e6, bytes / w / 1e6, bytes / cp / 1e6, bytes / mix / 1e6);
        }
        const float h2d = timed (s, [] (cudaStream_t st, void* p) { Ctx* c = (Ctx*) p; cudaMemcpyAsync (c->din, c->hin, c->n * 16, cudaMemcpyHostToDevice, st); }, &c);
        const float d2h = timed (s, [] (cudaStream_t st, void* p) { Ctx* c = (Ctx*) p; cudaMemcpyAsync (c->hout, c->dout, c->n * 16, cudaMemcpyDeviceToHost, st); }, &c);
        const float both = timed (s, [] (cudaStream_t st, void* p) {
            Ctx* c = (Ctx*) p;
            cudaEventRecord (c->ev, st);
            cudaStreamWaitEvent (c->s2, c->ev, 0);
            cudaMemcpyAsync (c->hout, c->dout, c->n * 16, cudaMemcpyDeviceToHost, c->s2);
            cudaMemcpyAsync (c->din, c->hin, c->n * 16, cudaMemcpyHostToDevice, st);
            cudaEventRecord (c->ev2, c->s2);
            cudaStreamWaitEvent (st, c->ev2, 0); }, &c);
        printf ("copy engines: H2D %6.1f GB/s | D2H %6.1f GB/s | both at once %6.1f GB/s each way\n", bytes / h2d / 1e6, bytes / d2h / 1e6, bytes / both / 1e6);
        cudaFreeHost (c.hin);
        cudaFreeHost (c.hout);
        cudaFree (c.din);
        cudaFree (c.dout);
    }
    printf ("%s\n", cudaGetErrorString (cudaGetLastError()));
    return 0;
}
