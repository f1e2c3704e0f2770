/*
 * abi_latency.c -- BASELINE.json configs[0] through the C ABI, without any host-language overhead: what the N-API shim
 * (wat-fft_b200/napi/watfft_napi.cc) would see per forward() call.
 *
 *   plan = wfb_plan_create(C2C, F32, SPLIT, n, batch)     <- createFFT*(size)        (index.js:69-91)
 *   per call: write the input views, wfb_exec(FORWARD)     <- input.set(); forward()  (benchmarks/lib/wat-contexts.js:125-129)
 *
 * Prints one JSON object: call latency (median / p10 / p99, microseconds) on the default path and with the zero-copy
 * path disabled, and the plan creation + destruction times (milliseconds, median of `reps`).
 * Build: make -C tools/native     Run: tools/native/abi_latency [n] [batch] [iters]
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "watfft_b200.h"

static double now_us(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec * 1e6 + t.tv_nsec * 1e-3; }
static int cmp(const void *a, const void *b) { double x = *(const double *)a, y = *(const double *)b; return x < y ? -1 : x > y; }

static void measure(wfb_plan *p, const float *re, const float *im, size_t bytes, int iters, double out[3]) {
    float *hre = (float *)wfb_host_in(p, 0), *him = (float *)wfb_host_in(p, 1);
    double *t = (double *)malloc(sizeof(double) * (size_t)iters);
    for (int i = 0; i < 200; i++) { memcpy(hre, re, bytes); memcpy(him, im, bytes); wfb_exec(p, WFB_FORWARD, WFB_EXEC_DEFAULT); }
    for (int i = 0; i < iters; i++) {
        const double t0 = now_us();
        memcpy(hre, re, bytes);
        memcpy(him, im, bytes);
        if (wfb_exec(p, WFB_FORWARD, WFB_EXEC_DEFAULT) != WFB_OK) { fprintf(stderr, "exec failed: %s\n", wfb_last_cuda_error()); exit(2); }
        t[i] = now_us() - t0;
    }
    qsort(t, (size_t)iters, sizeof(double), cmp);
    out[0] = t[iters / 2]; out[1] = t[iters / 10]; out[2] = t[(size_t)iters * 99 / 100];
    free(t);
}

int main(int argc, char **argv) {
    const int n = argc > 1 ? atoi(argv[1]) : 1024;
    const long batch = argc > 2 ? atol(argv[2]) : 1;
    const int iters = argc > 3 ? atoi(argv[3]) : 5000;
    int err = 0;
    if (wfb_require_b200(0) != WFB_OK) { printf("{\"error\": \"%s\"}\n", wfb_strerror(WFB_ERR_NO_DEVICE)); return 1; }
    wfb_plan *warm = wfb_plan_create(WFB_C2C, WFB_F32, WFB_SPLIT, n, batch, 0, &err);   /* context + module load, once per process */
    if (!warm) { printf("{\"error\": \"%s\"}\n", wfb_strerror(err)); return 1; }
    wfb_exec(warm, WFB_FORWARD, WFB_EXEC_DEFAULT);
    wfb_plan_destroy(warm);

    enum { REPS = 21 };
    double tc[REPS], td[REPS];
    for (int r = 0; r < REPS; r++) {
        double t0 = now_us();
        wfb_plan *p = wfb_plan_create(WFB_C2C, WFB_F32, WFB_SPLIT, n, batch, 0, &err);
        tc[r] = now_us() - t0;
        if (!p) { printf("{\"error\": \"%s\"}\n", wfb_strerror(err)); return 1; }
        t0 = now_us();
        wfb_plan_destroy(p);
        td[r] = now_us() - t0;
    }
    qsort(tc, REPS, sizeof(double), cmp);
    qsort(td, REPS, sizeof(double), cmp);

    const size_t bytes = sizeof(float) * (size_t)n * (size_t)batch;
    float *re = (float *)malloc(bytes), *im = (float *)malloc(bytes);
    unsigned s = 12345u;
    for (size_t i = 0; i < bytes / 4; i++) { s = s * 1664525u + 1013904223u; re[i] = (float)(s >> 8) / 8388608.0f - 1.0f; s = s * 1664525u + 1013904223u; im[i] = (float)(s >> 8) / 8388608.0f - 1.0f; }
    wfb_plan *p = wfb_plan_create(WFB_C2C, WFB_F32, WFB_SPLIT, n, batch, 0, &err);
    double dflt[3], copy[3];
    measure(p, re, im, bytes, iters, dflt);
    const int path = wfb_plan_last_path(p);
    wfb_plan_set_option(p, WFB_OPT_MAPPED_MAX_BYTES, 0);
    measure(p, re, im, bytes, iters / 4 > 200 ? iters / 4 : 200, copy);
    wfb_plan_destroy(p);
    printf("{\"n\": %d, \"batch\": %ld, \"iters\": %d, \"path\": %d, \"latency_us\": %.3f, \"latency_us_p10\": %.3f, \"latency_us_p99\": %.3f, "
           "\"copy_path_latency_us\": %.3f, \"plan_create_ms\": %.4f, \"plan_destroy_ms\": %.4f}\n",
           n, batch, iters, path, dflt[0], dflt[1], dflt[2], copy[0], tc[REPS / 2] * 1e-3, td[REPS / 2] * 1e-3);
    free(re); free(im);
    return 0;
}
