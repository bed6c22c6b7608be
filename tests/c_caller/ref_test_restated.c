/* TEST INFRASTRUCTURE.  The reference's own C test loop (/root/reference/test/test.c:9-172) restated against
 * include/chowdsp_fft.h and the product library: a plain C caller that really transforms on the GPU through the 11
 * drop-in symbols -- malloc'd and pre-allocated setups, real and complex, in place with an explicit work buffer, sizes
 * 2^5 .. 2^19 (or argv[1] .. argv[2]), SSE-layout and AVX-layout handles.
 *
 * The reference compares against pffft (not vendored, test/CMakeLists.txt:6).  Here the stand-in is the UNMODIFIED
 * reference itself (oracle/_ref/libchowdsp_fft_ref.so, built by oracle/Makefile), loaded with dlopen so that its
 * identically named symbols stay private to it.  Tolerances: the reference's own (test.c:12: 1e-6 * N / 8 per element)
 * plus the north star's relative L2 <= 1e-6 * log2 N.
 *
 * build: gcc -std=c11 -O1 -Iinclude tests/c_caller/ref_test_restated.c chowdsp_fft_b200/lib/libchowdsp_fft_b200.so -ldl -lm
 * run  : ./a.out [first_log2 last_log2 [path/to/libchowdsp_fft_ref.so]]
 */
#define _GNU_SOURCE
#include <chowdsp_fft.h>

#include <dlfcn.h>
#include <math.h>
#include <stdbool.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef void* (*new_setup_fn) (int, fft_transform_t, bool);
typedef void (*destroy_fn) (void*);
typedef void (*transform_fn) (void*, const float*, float*, float*, fft_direction_t);
typedef void* (*malloc_fn) (size_t);
typedef void (*free_fn) (void*);

static struct
{
    new_setup_fn new_setup;
    destroy_fn destroy;
    transform_fn transform;
    malloc_fn amalloc;
    free_fn afree;
} ref;

static int failures = 0;

static void compare (const float* r, const float* t, int n_floats, int N, const char* what)
{
    const float tol = 1.0e-6f * (float) n_floats / 8.0f; /* test.c:12 (N there is the float count) */
    double num = 0.0, den = 0.0;
    int bad = 0;
    for (int n = 0; n < n_floats; ++n)
    {
        const double d = (double) r[n] - (double) t[n];
        num += d * d;
        den += (double) r[n] * (double) r[n];
        if (! (fabsf (r[n] - t[n]) < tol))
            ++bad;
    }
    const double rel = sqrt (num / (den > 0.0 ? den : 1.0));
    const double rel_tol = 1.0e-6 * log2 ((double) N);
    if (bad != 0 || ! (rel <= rel_tol))
    {
        printf ("FAIL %s N=%d: %d elements beyond %g, relative L2 %.3g (limit %.3g)\n", what, N, bad, tol, rel, rel_tol);
        ++failures;
    }
}

static void run_case (int N, bool is_complex, bool use_avx, bool preallocate)
{
    const int nfl = is_complex ? 2 * N : N;
    const fft_transform_t kind = is_complex ? FFT_COMPLEX : FFT_REAL;
    float* data = (float*) aligned_malloc (sizeof (float) * (size_t) nfl);
    float* data_ref = (float*) ref.amalloc (sizeof (float) * (size_t) nfl);
    float* work_data = (float*) aligned_malloc (sizeof (float) * (size_t) nfl);
    float* work_data_ref = (float*) ref.amalloc (sizeof (float) * (size_t) nfl);
    if (data == NULL || data_ref == NULL || work_data == NULL || work_data_ref == NULL)
    {
        printf ("FAIL allocation N=%d\n", N);
        ++failures;
        return;
    }
    for (int i = 0; i < N; ++i)
    {
        if (is_complex)
        {
            data[i * 2] = sinf (3.14f * (100.0f / 48000.0f) * (float) i);
            data[i * 2 + 1] = cosf (3.14f * (100.0f / 48000.0f) * (float) i);
        }
        else
            data[i] = sinf (3.14f * (100.0f / 48000.0f) * (float) i);
    }
    memcpy (data_ref, data, (size_t) nfl * sizeof (float));

    void* fft_setup;
    void* prealloc = NULL;
    if (preallocate)
    {
        const size_t bytes_required = fft_bytes_required (N, kind, use_avx);
        prealloc = aligned_malloc (bytes_required);
        fft_setup = fft_new_setup_preallocated (N, kind, prealloc, use_avx);
    }
    else
        fft_setup = fft_new_setup (N, kind, use_avx);
    void* ref_setup = ref.new_setup (N, kind, use_avx);
    if (fft_setup == NULL || ref_setup == NULL || ref_setup == (void*) 1)
    {
        printf ("FAIL setup N=%d complex=%d avx=%d prealloc=%d (%p / %p)\n", N, (int) is_complex, (int) use_avx, (int) preallocate, fft_setup, ref_setup);
        ++failures;
        return;
    }

    fft_transform (fft_setup, data, data, work_data, FFT_FORWARD);
    ref.transform (ref_setup, data_ref, data_ref, work_data_ref, FFT_FORWARD);
    compare (data_ref, data, nfl, N, is_complex ? "complex forward" : "real forward");

    fft_transform (fft_setup, data, data, work_data, FFT_BACKWARD);
    ref.transform (ref_setup, data_ref, data_ref, work_data_ref, FFT_BACKWARD);
    const float norm_gain = 1.0f / (float) N;
    for (int n = 0; n < nfl; ++n)
    {
        data[n] *= norm_gain;
        data_ref[n] *= norm_gain;
    }
    compare (data_ref, data, nfl, N, is_complex ? "complex round trip" : "real round trip");

    if (preallocate)
        aligned_free (prealloc); /* the caller owns the block; no fft_destroy_setup (chowdsp_fft.h:98-113) */
    else
        fft_destroy_setup (fft_setup);
    ref.destroy (ref_setup);
    aligned_free (data);
    ref.afree (data_ref);
    aligned_free (work_data);
    ref.afree (work_data_ref);
}

int main (int argc, char** argv)
{
    const int first = argc > 2 ? atoi (argv[1]) : 5, last = argc > 2 ? atoi (argv[2]) : 19;
    const char* ref_path = argc > 3 ? argv[3] : "oracle/_ref/libchowdsp_fft_ref.so";
    void* h = dlopen (ref_path, RTLD_NOW | RTLD_LOCAL | RTLD_DEEPBIND);
    if (h == NULL)
    {
        printf ("cannot load the reference build %s: %s\n", ref_path, dlerror());
        return 2;
    }
    ref.new_setup = (new_setup_fn) dlsym (h, "fft_new_setup");
    ref.destroy = (destroy_fn) dlsym (h, "fft_destroy_setup");
    ref.transform = (transform_fn) dlsym (h, "fft_transform");
    ref.amalloc = (malloc_fn) dlsym (h, "aligned_malloc");
    ref.afree = (free_fn) dlsym (h, "aligned_free");
    if (ref.new_setup == NULL || ref.destroy == NULL || ref.transform == NULL || ref.amalloc == NULL || ref.afree == NULL)
    {
        printf ("the reference build lacks one of the drop-in symbols\n");
        return 2;
    }
    if (ref.new_setup == (new_setup_fn) fft_new_setup)
    {
        printf ("symbol interposition: the stand-in resolves to the product\n");
        return 2;
    }
    int cases = 0;
    for (int mode = 0; mode < 3; ++mode) /* test.c:139-166: SSE-layout handles, AVX-layout handles, pre-allocated */
    {
        const bool use_avx = mode == 1, prealloc = mode == 2;
        printf ("%s\n", mode == 0 ? "Running SSE-layout tests" : mode == 1 ? "Running AVX-layout tests" : "Running pre-allocated tests");
        for (int i = first; i <= last; ++i)
        {
            run_case (1 << i, true, use_avx, prealloc);
            run_case (1 << i, false, use_avx, prealloc);
            cases += 2;
        }
    }
    printf ("%d cases, %d failures\n", cases, failures);
    if (failures == 0)
        printf ("Testing complete!\n");
    return failures == 0 ? 0 : 1;
}
