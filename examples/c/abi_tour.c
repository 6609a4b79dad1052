/* A C99 caller of liblele_b200.so: what a binding in any language does, in the language the ABI is written in.
 *
 *   gcc -std=c99 -I include examples/c/abi_tour.c -L lele_b200 -llele_b200 -Wl,-rpath,$PWD/lele_b200 -o abi_tour
 *
 * Host-only entry points (window, filterbank, frame arithmetic) run anywhere.  Everything else needs a B200: without a
 * device the program says so and exits 0 -- there is no CPU fallback behind this header.  With a device it runs one
 * operator (layer_norm, lele::kernels::layer_norm, src/kernels/norm.rs:226) through the explicit-copy path. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "lele_b200.h"

#define CHECK(call)                                                                      \
    do {                                                                                 \
        int rc_ = (call);                                                                \
        if (rc_ != 0) {                                                                  \
            fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, lele_b200_last_error()); \
            return 1;                                                                    \
        }                                                                                \
    } while (0)

int main(void) {
    /* lele::features::hann_window / mel_filterbank / SenseVoiceFrontend frame counts: host arithmetic */
    float win[4];
    CHECK(lele_b200_hann_window(4, win));
    printf("hann(4) = %.2f %.2f %.2f %.2f\n", win[0], win[1], win[2], win[3]);
    float* fb = (float*)malloc(sizeof(float) * 80 * 201);
    if (!fb) return 1;
    CHECK(lele_b200_mel_filterbank(16000.0f, 400, 80, 20.0f, 8000.0f, fb));
    float peak = 0.0f;
    for (int i = 0; i < 80 * 201; ++i) peak = fb[i] > peak ? fb[i] : peak;
    printf("mel filterbank 80 x 201, peak weight %.3f\n", peak);
    free(fb);
    printf("16 s clip: %d frames -> %d LFR rows\n", lele_b200_frontend_num_frames(256000), lele_b200_frontend_out_rows(256000));

    if (lele_b200_device_count() <= 0) {
        printf("no CUDA device: compute entry points are unavailable (by design there is no CPU fallback)\n");
        return 0;
    }

    lele_b200_ctx* ctx = NULL;
    CHECK(lele_b200_ctx_create(0, NULL, &ctx));
    const float x[6] = {1.0f, 2.0f, 3.0f, -1.0f, 0.0f, 1.0f}, gamma[3] = {1.0f, 1.0f, 1.0f}, beta[3] = {0.0f, 0.0f, 0.0f};
    float y[6];
    void *dx = NULL, *dg = NULL, *db = NULL, *dy = NULL;
    CHECK(lele_b200_malloc(ctx, sizeof x, &dx));
    CHECK(lele_b200_malloc(ctx, sizeof gamma, &dg));
    CHECK(lele_b200_malloc(ctx, sizeof beta, &db));
    CHECK(lele_b200_malloc(ctx, sizeof y, &dy));
    CHECK(lele_b200_h2d(ctx, dx, x, sizeof x));
    CHECK(lele_b200_h2d(ctx, dg, gamma, sizeof gamma));
    CHECK(lele_b200_h2d(ctx, db, beta, sizeof beta));
    CHECK(lele_b200_layer_norm(ctx, (const float*)dx, (const float*)dg, (const float*)db, 2, 3, 1e-5f, (float*)dy));
    CHECK(lele_b200_d2h(ctx, y, dy, sizeof y));
    CHECK(lele_b200_sync(ctx));
    printf("layer_norm([1,2,3]) = %.6f %.6f %.6f   (tests/verify_operators.rs:34 expects -1.2247356 0 1.2247356)\n", y[0], y[1], y[2]);
    int ok = fabsf(y[0] + 1.2247356f) < 1e-4f && fabsf(y[1]) < 1e-4f && fabsf(y[2] - 1.2247356f) < 1e-4f;
    CHECK(lele_b200_free(ctx, dx));
    CHECK(lele_b200_free(ctx, dg));
    CHECK(lele_b200_free(ctx, db));
    CHECK(lele_b200_free(ctx, dy));
    printf("kernel launches issued by this context: %llu\n", lele_b200_launch_count(ctx));
    CHECK(lele_b200_ctx_destroy(ctx));
    return ok ? 0 : 2;
}
