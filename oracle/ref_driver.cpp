// TEST INFRASTRUCTURE ONLY (oracle/): thread-per-core batch driver over the UNMODIFIED reference API.
// Linked into oracle/_ref/libchowdsp_fft_ref.so next to the reference's own objects.  The reference is
// single-threaded per transform and its setup is read-only/shareable (chowdsp_fft.h:87-91), so the CPU
// baseline for a batch is "one thread per core, each looping over its contiguous slice of the batch"
// with a private work buffer.  Only tests/, bench.py's cpu_baseline /
// --impl reference legs and __graft_entry__.smoke() may load this.
#include <chowdsp_fft.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstring>
#include <thread>
#include <vector>

#if defined(__linux__)
#include <pthread.h>
#include <sched.h>
#endif

using namespace chowdsp::fft;

namespace
{
void pin_to_core (int core)
{
#if defined(__linux__)
    cpu_set_t allowed;
    CPU_ZERO (&allowed);
    if (sched_getaffinity (0, sizeof (allowed), &allowed) != 0)
        return;
    std::vector<int> cpus;
    for (int c = 0; c < CPU_SETSIZE; ++c)
        if (CPU_ISSET (c, &allowed))
            cpus.push_back (c);
    if (cpus.empty())
        return;
    cpu_set_t one;
    CPU_ZERO (&one);
    CPU_SET (cpus[(size_t) core % cpus.size()], &one);
    pthread_setaffinity_np (pthread_self(), sizeof (one), &one);
#else
    (void) core;
#endif
}

template <typename Fn>
double run_threads (int nthreads, long batch, Fn&& body)
{
    nthreads = (int) std::max<long> (1, std::min<long> (nthreads, batch));
    std::vector<std::thread> pool;
    const auto t0 = std::chrono::steady_clock::now();
    for (int t = 0; t < nthreads; ++t)
    {
        const long lo = batch * t / nthreads, hi = batch * (t + 1) / nthreads;
        pool.emplace_back ([=, &body]
                           {
                               pin_to_core (t);
                               body (lo, hi);
                           });
    }
    for (auto& th : pool)
        th.join();
    return std::chrono::duration<double> (std::chrono::steady_clock::now() - t0).count();
}
} // namespace

extern "C"
{
// batch of transforms through fft_transform / fft_transform_unordered; strides in floats.
// Returns wall seconds of the threaded region (buffers and setup are made outside it).
double ref_transform_batched (int N, int is_complex, int use_avx, int backward, int ordered, const float* in, float* out, long batch, long in_stride, long out_stride, int nthreads)
{
    void* setup = fft_new_setup (N, is_complex ? FFT_COMPLEX : FFT_REAL, use_avx != 0);
    if (setup == nullptr || setup == (void*) 1)
        return -1.0;
    const size_t nfloats = (size_t) N * (is_complex ? 2 : 1);
    const double secs = run_threads (nthreads, batch, [&] (long lo, long hi)
                                     {
                                         auto* work = (float*) aligned_malloc (nfloats * sizeof (float));
                                         for (long b = lo; b < hi; ++b)
                                         {
                                             if (ordered)
                                                 fft_transform (setup, in + b * in_stride, out + b * out_stride, work, backward ? FFT_BACKWARD : FFT_FORWARD);
                                             else
                                                 fft_transform_unordered (setup, in + b * in_stride, out + b * out_stride, work, backward ? FFT_BACKWARD : FFT_FORWARD);
                                         }
                                         aligned_free (work);
                                     });
    fft_destroy_setup (setup);
    return secs;
}

// The same loop `reps` times inside ONE set of pinned threads: plan, work buffers and threads are created before the timed
// region, which starts when every thread is ready (bench.py's CPU arms: no thread spawn or fft_new_setup inside the timing).
double ref_transform_batched_reps (int N, int is_complex, int use_avx, int backward, int ordered, const float* in, float* out, long batch, long in_stride, long out_stride, int nthreads, int reps)
{
    void* setup = fft_new_setup (N, is_complex ? FFT_COMPLEX : FFT_REAL, use_avx != 0);
    if (setup == nullptr || setup == (void*) 1)
        return -1.0;
    const size_t nfloats = (size_t) N * (is_complex ? 2 : 1);
    nthreads = (int) std::max<long> (1, std::min<long> (nthreads, batch));
    std::atomic<int> ready { 0 }, go { 0 };
    std::vector<std::thread> pool;
    for (int t = 0; t < nthreads; ++t)
    {
        const long lo = batch * t / nthreads, hi = batch * (t + 1) / nthreads;
        pool.emplace_back ([=, &ready, &go]
                           {
                               pin_to_core (t);
                               auto* work = (float*) aligned_malloc (nfloats * sizeof (float));
                               ready.fetch_add (1);
                               while (go.load (std::memory_order_acquire) == 0)
                                   std::this_thread::yield();
                               for (int r = 0; r < reps; ++r)
                                   for (long b = lo; b < hi; ++b)
                                   {
                                       if (ordered)
                                           fft_transform (setup, in + b * in_stride, out + b * out_stride, work, backward ? FFT_BACKWARD : FFT_FORWARD);
                                       else
                                           fft_transform_unordered (setup, in + b * in_stride, out + b * out_stride, work, backward ? FFT_BACKWARD : FFT_FORWARD);
                                   }
                               aligned_free (work);
                           });
    }
    while (ready.load() < nthreads)
        std::this_thread::yield();
    const auto t0 = std::chrono::steady_clock::now();
    go.store (1, std::memory_order_release);
    for (auto& th : pool)
        th.join();
    const double secs = std::chrono::duration<double> (std::chrono::steady_clock::now() - t0).count();
    fft_destroy_setup (setup);
    return secs;
}

// batch of ab += a*b*scaling through fft_convolve_unordered (strides in floats; 0 = shared operand).
double ref_convolve_batched (int N, int is_complex, int use_avx, const float* a, const float* b, float* ab, long batch, long a_stride, long b_stride, long ab_stride, float scaling, int nthreads)
{
    void* setup = fft_new_setup (N, is_complex ? FFT_COMPLEX : FFT_REAL, use_avx != 0);
    if (setup == nullptr || setup == (void*) 1)
        return -1.0;
    const double secs = run_threads (nthreads, batch, [&] (long lo, long hi)
                                     {
                                         for (long i = lo; i < hi; ++i)
                                             fft_convolve_unordered (setup, a + i * a_stride, b + i * b_stride, ab + i * ab_stride, scaling);
                                     });
    fft_destroy_setup (setup);
    return secs;
}

// Uniform partitioned overlap-save convolution (BASELINE config 4) through the reference API, one
// channel per loop iteration: per block 1 unordered R2C + P fft_convolve_unordered + 1 unordered C2R.
//   x   [channels][blocks*B]   input samples, B = N/2
//   h   [channels][P][N]       per-channel IR partitions, already transformed (unordered)
//   fdl [channels][P][N]       frequency-delay line (ring), zero-initialised by the caller
//   y   [channels][blocks*B]   output samples
// scaling (normally 1/N) is folded into the MAC.  Returns wall seconds.
double ref_partitioned_convolve (int N, int P, int use_avx, const float* x, const float* h, float* fdl, float* y, long channels, int blocks, int first_block, int nthreads)
{
    void* setup = fft_new_setup (N, FFT_REAL, use_avx != 0);
    if (setup == nullptr || setup == (void*) 1)
        return -1.0;
    const int B = N / 2;
    const float scaling = 1.0f / (float) N;
    const double secs = run_threads (nthreads, channels, [&] (long lo, long hi)
                                     {
                                         auto* work = (float*) aligned_malloc ((size_t) N * sizeof (float));
                                         auto* win = (float*) aligned_malloc ((size_t) N * sizeof (float));
                                         auto* acc = (float*) aligned_malloc ((size_t) N * sizeof (float));
                                         for (long c = lo; c < hi; ++c)
                                         {
                                             const float* xc = x + c * (long) blocks * B;
                                             float* yc = y + c * (long) blocks * B;
                                             for (int t = first_block; t < blocks; ++t)
                                             {
                                                 // overlap-save window: previous block then the new block
                                                 if (t == 0)
                                                     std::memset (win, 0, (size_t) B * sizeof (float));
                                                 else
                                                     std::memcpy (win, xc + (long) (t - 1) * B, (size_t) B * sizeof (float));
                                                 std::memcpy (win + B, xc + (long) t * B, (size_t) B * sizeof (float));
                                                 float* slot = fdl + (c * P + (t % P)) * (long) N;
                                                 fft_transform_unordered (setup, win, slot, work, FFT_FORWARD);
                                                 std::memset (acc, 0, (size_t) N * sizeof (float));
                                                 for (int p = 0; p < P; ++p)
                                                 {
                                                     if (t - p < 0)
                                                         break;
                                                     const float* X = fdl + (c * P + ((t - p) % P)) * (long) N;
                                                     const float* H = h + (c * P + p) * (long) N;
                                                     fft_convolve_unordered (setup, X, H, acc, scaling);
                                                 }
                                                 fft_transform_unordered (setup, acc, win, work, FFT_BACKWARD);
                                                 std::memcpy (yc + (long) t * B, win + B, (size_t) B * sizeof (float));
                                             }
                                         }
                                         aligned_free (work);
                                         aligned_free (win);
                                         aligned_free (acc);
                                     });
    fft_destroy_setup (setup);
    return secs;
}

int ref_hardware_threads()
{
    return (int) std::max (1u, std::thread::hardware_concurrency());
}
}
