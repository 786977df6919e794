/* The C ABI used from plain C, without Python: flag one small dump and print a digest.
 *
 *   gcc -std=c99 -I include tests/c/flagger_smoke.c -o flagger_smoke \
 *       -L katsdpsigproc_b200/_lib -lksp_b200 -Wl,-rpath,$PWD/katsdpsigproc_b200/_lib -lm
 *
 * tests/test_c_example.py builds it (CPU) and runs it (GPU), comparing the digest with the
 * oracle on the same synthetic dump. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "ksp_b200.h"

#define CHECK(call)                                                              \
    do {                                                                         \
        int rc_ = (call);                                                        \
        if (rc_ != 0) {                                                          \
            fprintf(stderr, "%s failed: %s\n", #call, ksp_error_string(rc_));    \
            return 1;                                                            \
        }                                                                        \
    } while (0)

/* xorshift: the Python side regenerates the same dump */
static uint32_t rng_state = 12345u;
static float next_uniform(void)
{
    rng_state ^= rng_state << 13;
    rng_state ^= rng_state >> 17;
    rng_state ^= rng_state << 5;
    return (float) (rng_state >> 8) * (1.0f / 16777216.0f);
}

int main(int argc, char **argv)
{
    const int64_t channels = argc > 1 ? atoll(argv[1]) : 1024;
    const int64_t baselines = argc > 2 ? atoll(argv[2]) : 40;
    const size_t n = (size_t) channels * (size_t) baselines;
    float *vis = (float *) malloc(n * 2 * sizeof(float));
    uint8_t *flags = (uint8_t *) malloc(n);
    float *noise = (float *) malloc((size_t) baselines * sizeof(float));
    if (!vis || !flags || !noise) return 2;
    for (size_t i = 0; i < n; i++) {
        float re = next_uniform() - 0.5f, im = next_uniform() - 0.5f;
        if (next_uniform() < 1.0f / 64.0f) re += 20.0f;       /* narrow-band interference */
        vis[2 * i] = re;
        vis[2 * i + 1] = im;
    }

    int count = 0;
    CHECK(ksp_device_count(&count));
    if (count == 0) {
        fprintf(stderr, "no CUDA device\n");
        return 3;
    }
    CHECK(ksp_device_set(0));
    void *stream = NULL, *d_vis = NULL, *d_flags = NULL, *d_noise = NULL, *d_scratch = NULL;
    CHECK(ksp_stream_create(&stream));

    ksp_flagger_params prm;
    memset(&prm, 0, sizeof(prm));
    prm.channels = channels;
    prm.baselines = baselines;
    prm.vis_stride = baselines;
    prm.flags_stride = baselines;
    prm.width = 13;
    prm.flag_mode = KSP_FLAGS_NONE;
    prm.abs_mode = KSP_ABS_NUMPY;
    prm.n_windows = 7;
    prm.flag_value = 1;
    prm.n_sigma = 11.0;
    for (int w = 0; w < prm.n_windows; w++) prm.scales[w] = pow(1.2, -w);
    const size_t scratch_bytes = ksp_flagger_scratch_bytes(&prm);

    CHECK(ksp_malloc(&d_vis, n * 8));
    CHECK(ksp_malloc(&d_flags, n));
    CHECK(ksp_malloc(&d_noise, (size_t) baselines * 4));
    CHECK(ksp_malloc(&d_scratch, scratch_bytes));
    CHECK(ksp_memcpy_async(d_vis, vis, n * 8, 1 /* host to device */, stream));
    CHECK(ksp_flagger(stream, &prm, d_vis, NULL, (float *) d_noise, (uint8_t *) d_flags, d_scratch,
                      scratch_bytes));
    CHECK(ksp_memcpy_async(flags, d_flags, n, 2 /* device to host */, stream));
    CHECK(ksp_memcpy_async(noise, d_noise, (size_t) baselines * 4, 2 /* device to host */, stream));
    CHECK(ksp_stream_synchronize(stream));

    unsigned long long flagged = 0, digest = 1469598103934665603ull;   /* FNV-1a over the flags */
    for (size_t i = 0; i < n; i++) {
        flagged += flags[i];
        digest = (digest ^ flags[i]) * 1099511628211ull;
    }
    double noise_sum = 0.0;
    for (int64_t b = 0; b < baselines; b++) noise_sum += noise[b];
    printf("flagged %llu digest %llu noise_sum %.9g\n", flagged, digest, noise_sum);

    CHECK(ksp_free(d_scratch));
    CHECK(ksp_free(d_noise));
    CHECK(ksp_free(d_flags));
    CHECK(ksp_free(d_vis));
    CHECK(ksp_stream_destroy(stream));
    free(noise);
    free(flags);
    free(vis);
    return 0;
}
