/* TEST / BENCH INFRASTRUCTURE.  The reference's own bench loop (/root/reference/bench/bench.cpp:75-110: one real transform of
 * N samples, forward + backward in place on aligned_malloc buffers, M round trips) as a compiled C caller of the product
 * library -- the latency of the synchronous drop-in calls without any Python binding overhead.
 * Prints: us_per_round_trip <value> (best of 5 blocks of M round trips), plus the same for the in-place scaled check.
 * build: gcc -std=c11 -O2 -Iinclude tests/c_caller/latency_bench.c chowdsp_fft_b200/lib/libchowdsp_fft_b200.so -lm
 */
#define _POSIX_C_SOURCE 200809L
#include <chowdsp_fft.h>

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <time.h>

static double now_s (void)
{
    struct timespec ts;
    clock_gettime (CLOCK_MONOTONIC, &ts);
    return (double) ts.tv_sec + 1e-9 * (double) ts.tv_nsec;
}

int main (int argc, char** argv)
{
    const int N = argc > 1 ? atoi (argv[1]) : 1024, M = argc > 2 ? atoi (argv[2]) : 2000;
    float* data = (float*) aligned_malloc (sizeof (float) * (size_t) N);
    float* work = (float*) aligned_malloc (sizeof (float) * (size_t) N);
    float* keep = (float*) malloc (sizeof (float) * (size_t) N);
    void* setup = fft_new_setup (N, FFT_REAL, true);
    if (data == NULL || work == NULL || keep == NULL || setup == NULL)
    {
        printf ("setup failed\n");
        return 2;
    }
    for (int i = 0; i < N; ++i)
        keep[i] = data[i] = sinf (3.14f * (100.0f / 48000.0f) * (float) i);
    /* warm-up + correctness: one round trip returns N * x */
    fft_transform (setup, data, data, work, FFT_FORWARD);
    fft_transform (setup, data, data, work, FFT_BACKWARD);
    double err = 0.0, ref = 0.0;
    for (int i = 0; i < N; ++i)
    {
        const double d = (double) data[i] / N - (double) keep[i];
        err += d * d;
        ref += (double) keep[i] * (double) keep[i];
        data[i] = keep[i];
    }
    double best = 1e30;
    for (int block = 0; block < 5; ++block)
    {
        const double t0 = now_s();
        for (int m = 0; m < M; ++m)
        {
            fft_transform (setup, data, data, work, FFT_FORWARD);
            fft_transform (setup, data, data, work, FFT_BACKWARD);
            if ((m & 7) == 7) /* keep the values bounded: the transforms are unnormalised */
                for (int i = 0; i < N; ++i)
                    data[i] = keep[i];
        }
        const double dt = (now_s() - t0) / M;
        if (dt < best)
            best = dt;
    }
    printf ("N %d round_trips %d us_per_round_trip %.3f us_per_call %.3f round_trip_rel_l2 %.3g\n", N, M, best * 1e6, best * 0.5e6, sqrt (err / ref));
    fft_destroy_setup (setup);
    aligned_free (data);
    aligned_free (work);
    free (keep);
    return sqrt (err / ref) < 1e-5 ? 0 : 1;
}
