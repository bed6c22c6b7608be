/*
 * chowdsp_fft drop-in C API, implemented by libchowdsp_fft_b200.so on NVIDIA B200 (sm_100a).
 *
 * Every declaration below has the same name, signature, enum values and calling convention as the
 * reference's public header (/root/reference/chowdsp_fft.h, line cited per entry), so code written
 * against chowdsp_fft re-links unchanged.  The comments are ours and describe the GPU behaviour.
 * Batched / stream-ordered extensions live in chowdsp_fft_b200.h.
 *
 * Pointer policy (new; the reference only knows host memory):
 *   - device and managed pointers are used in place;
 *   - host pointers work too: memory from aligned_malloc() is pinned and mapped, small transforms
 *     run on it zero-copy, larger ones and ordinary malloc'd memory are staged through the device;
 *   - every function in THIS header has returned its result when it returns (host-synchronous),
 *     exactly like the reference.  Use chowdsp_fft_b200.h for asynchronous, batched work.
 * There is no CPU implementation behind this API: without a usable CUDA device fft_new_setup returns
 * NULL (fft_b200_last_error() says why) and nothing else may be called.
 */
#pragma once

#ifdef __cplusplus
#include <cstddef>

extern "C"
{
namespace chowdsp::fft
{
#else
#include <stdbool.h>
#include <stddef.h>
#endif

/* reference chowdsp_fft.h:64-68 */
typedef enum
{
    FFT_FORWARD,
    FFT_BACKWARD
} fft_direction_t;

/* reference chowdsp_fft.h:71-75 */
typedef enum
{
    FFT_REAL,
    FFT_COMPLEX
} fft_transform_t;

#ifdef __cplusplus
#define CHOWDSP_FFT_DEFAULT_TRUE = true
#else
#define CHOWDSP_FFT_DEFAULT_TRUE
#endif

/* reference chowdsp_fft.h:81.  Bytes the caller must provide to fft_new_setup_preallocated.  Never
   less than the reference's figure (8N+96 complex / 4N+96 real), so existing arenas stay big enough;
   only the small host-side handle lives there, the twiddle tables are device memory owned by a
   process-wide plan cache. */
size_t fft_bytes_required (int N, fft_transform_t transform, bool use_avx_if_available CHOWDSP_FFT_DEFAULT_TRUE);

/* reference chowdsp_fft.h:92.  Plan for power-of-two N: real 32 <= N <= 2^28, complex 16 <= N <= 2^28.
   Up to 32768 real / 16384 complex points a transform is ONE kernel; larger sizes run as a two- or
   three-pass four-step transform.  The handle is immutable and may be shared between threads.
   use_avx_if_available selects which of the reference's two "unordered" layouts the plan speaks:
   true  -> the 8-lane (AVX) layout when N % 128 == 0 (real) / N % 64 == 0 (complex), else 4-lane;
   false -> always the 4-lane (SSE/NEON) layout.  Returns NULL for an unsupported N or without a GPU. */
void* fft_new_setup (int N, fft_transform_t transform, bool use_avx_if_available CHOWDSP_FFT_DEFAULT_TRUE);

/* reference chowdsp_fft.h:114.  Same, with the handle placed in caller memory `data`
   (>= fft_bytes_required bytes, 8-byte aligned or better).  Such a handle needs no
   fft_destroy_setup; the caller frees `data`. */
void* fft_new_setup_preallocated (int N, fft_transform_t transform, void* data, bool use_avx_if_available CHOWDSP_FFT_DEFAULT_TRUE);

/* reference chowdsp_fft.h:119 */
void fft_destroy_setup (void*);

/* reference chowdsp_fft.h:122.  32 when the plan uses the 8-lane unordered layout, else 16.  Data
   buffers must be aligned to at least this many bytes, as with the reference. */
int fft_simd_width_bytes (void* setup);

/* reference chowdsp_fft.h:138.  Ordered transform.  Complex: interleaved (re,im), natural bin order.
   Real: forward output / backward input is the packed half spectrum
   [X0.re, X(N/2).re, X1.re, X1.im, ..., X(N/2-1).re, X(N/2-1).im].  Unscaled:
   BACKWARD(FORWARD(x)) = N x.  input and output may alias.  `work` is accepted for compatibility and
   ignored (the scratch is on-chip shared memory). */
void fft_transform (void* setup, const float* input, float* output, float* work, fft_direction_t direction);

/* reference chowdsp_fft.h:145.  Same, but the frequency-domain side uses the plan's unordered layout
   (bit-compatible, slot for slot, with the reference's for the same N and SIMD width). */
void fft_transform_unordered (void* setup, const float* input, float* output, float* work, fft_direction_t direction);

/* reference chowdsp_fft.h:154.  dft_ab += (dft_a * dft_b) * scaling on unordered spectra.  For a
   real plan the DC and Nyquist slots are two independent real products.  The pointers may alias. */
void fft_convolve_unordered (void* setup, const float* dft_a, const float* dft_b, float* dft_ab, float scaling);

/* reference chowdsp_fft.h:160.  ab[i] = a[i] + b[i], i < N; N a multiple of 2 * (SIMD width in floats). */
void fft_accumulate (void* setup, const float* a, const float* b, float* ab, int N);

/* reference chowdsp_fft.h:162-163.  64-byte aligned, page-locked, device-mapped host memory (falls
   back to ordinary 64-byte aligned host memory when no CUDA device is usable). */
void* aligned_malloc (size_t nb_bytes);
void aligned_free (void*);

#undef CHOWDSP_FFT_DEFAULT_TRUE

#ifdef __cplusplus
}
} // namespace chowdsp::fft
#endif
