// Unordered-domain elementwise kernels: the frequency-domain multiply-accumulate behind
// fft_convolve_unordered and the plain sum behind fft_accumulate.  Pure HBM streaming.
#pragma once
#include "fft_kernels.cuh"

namespace cfb
{
// ---------------------------------------------------------------------------------------------
// unordered-domain elementwise kernels
// ---------------------------------------------------------------------------------------------
struct ConvArgs
{
    const float* a;
    const float* b;
    float* ab;
    long long a_stride, b_stride, ab_stride; // floats between consecutive spectra (0 = shared operand)
    int nfloats;                             // floats per spectrum (N real, 2N complex)
    int batch;
    int logW;
    int is_real;                             // DC / Nyquist are two real products (avx:1974-1978)
    float scaling;
};

// ab += a * b * scaling on (W re | W im) vector pairs -- pffft_convolve_internal,
// /root/reference/simd/chowdsp_fft_impl_avx.cpp:1937-1979.  One thread per 4 re + 4 im floats.
__global__ void __launch_bounds__ (256) convolve_kernel (const ConvArgs p)
{
    const int W = 1 << p.logW;
    const int quads = p.nfloats >> 3; // work items per spectrum
    const long long total = (long long) quads * p.batch;
    for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long) gridDim.x * blockDim.x)
    {
        const long long x = i / quads;
        const int q = (int) (i - x * quads);
        // W = 8: two quads per 16-float chunk;  W = 4: one quad per 8-float chunk
        const int off = p.logW == 3 ? ((q >> 1) * 16 + (q & 1) * 4) : q * 8;
        // plain loads: a, b and ab may alias (chowdsp_fft.h:153)
        const float4 ar = *reinterpret_cast<const float4*> (p.a + x * p.a_stride + off);
        const float4 ai = *reinterpret_cast<const float4*> (p.a + x * p.a_stride + off + W);
        const float4 br = *reinterpret_cast<const float4*> (p.b + x * p.b_stride + off);
        const float4 bi = *reinterpret_cast<const float4*> (p.b + x * p.b_stride + off + W);
        float4* pr = reinterpret_cast<float4*> (p.ab + x * p.ab_stride + off);
        float4* pi = reinterpret_cast<float4*> (p.ab + x * p.ab_stride + off + W);
        float4 cr = *pr, ci = *pi;
        const float s = p.scaling;
        const float dc = cr.x + (ar.x * br.x) * s, ny = ci.x + (ai.x * bi.x) * s;
        cr.x += (ar.x * br.x - ai.x * bi.x) * s; ci.x += (ar.x * bi.x + ai.x * br.x) * s;
        cr.y += (ar.y * br.y - ai.y * bi.y) * s; ci.y += (ar.y * bi.y + ai.y * br.y) * s;
        cr.z += (ar.z * br.z - ai.z * bi.z) * s; ci.z += (ar.z * bi.z + ai.z * br.z) * s;
        cr.w += (ar.w * br.w - ai.w * bi.w) * s; ci.w += (ar.w * bi.w + ai.w * br.w) * s;
        if (p.is_real && q == 0)
        {
            cr.x = dc;
            ci.x = ny;
        }
        *pr = cr;
        *pi = ci;
    }
}

// ab = a + b -- fft_accumulate_internal, avx:1981-1994
__global__ void __launch_bounds__ (256) accumulate_kernel (const float* a, const float* b, float* ab, long long n4)
{
    for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long) gridDim.x * blockDim.x)
    {
        const float4 x = reinterpret_cast<const float4*> (a)[i], y = reinterpret_cast<const float4*> (b)[i];
        reinterpret_cast<float4*> (ab)[i] = make_float4 (x.x + y.x, x.y + y.y, x.z + y.z, x.w + y.w);
    }
}

// JUCE convention for real spectra when the caller wants the redundant negative frequencies as well
// (chowdsp_fft_juce.cpp:58-61): bin M + i = conj (bin M - i), i = 1 .. M-1, rows of 2 M complex bins
__global__ void __launch_bounds__ (256) juce_mirror_kernel (float* data, long long stride, int M, long long total)
{
    for (long long t = (long long) blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long) gridDim.x * blockDim.x)
    {
        const long long row = t / (M - 1);
        const int i = 1 + (int) (t - row * (M - 1));
        float2* d = reinterpret_cast<float2*> (data + row * stride);
        const float2 v = d[M - i];
        d[M + i] = make_float2 (v.x, -v.y);
    }
}

} // namespace cfb
