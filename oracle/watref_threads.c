/*
 * watref_threads.c -- pthread driver around the transpiled reference modules
 * (TEST INFRASTRUCTURE; links into oracle/_ref/libwatref.so).
 *
 * The reference is single-threaded: one WebAssembly.Instance with private
 * linear memory per context (index.js:13-18).  The CPU baseline named by
 * BASELINE.json ("reference WASM on all host cores via worker_threads") is
 * restated here as one private module memory per pthread; each thread loops
 *     memcpy(input row) -> transform(n)
 * over its share of the rows exactly as bench() does in
 * benchmarks/lib/wat-contexts.js:125-129, with precompute done once per thread.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef void (*wat_fn)(uint8_t *, uint32_t);

typedef struct {
    wat_fn precompute, run;
    uint32_t pages, n;
    const uint8_t *in0; size_t in0_bytes, in0_stride; uint32_t dst0;   /* plane 0 -> memory offset dst0 */
    const uint8_t *in1; size_t in1_bytes, in1_stride; uint32_t dst1;   /* optional plane 1 */
    uint8_t *out0; size_t out0_bytes, out0_stride; uint32_t src0;      /* optional copy-out */
    uint8_t *out1; size_t out1_bytes, out1_stride; uint32_t src1;
    long row_begin, row_end, reps;
    double seconds;
} job_t;

static double now_s(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }

static void *worker(void *arg) {
    job_t *j = (job_t *)arg;
    uint8_t *mem = (uint8_t *)aligned_alloc(64, (size_t)j->pages * 65536);
    memset(mem, 0, (size_t)j->pages * 65536);
    j->precompute(mem, j->n);
    double t0 = now_s();
    for (long rep = 0; rep < j->reps; rep++) {
        for (long r = j->row_begin; r < j->row_end; r++) {
            memcpy(mem + j->dst0, j->in0 + (size_t)r * j->in0_stride, j->in0_bytes);
            if (j->in1) memcpy(mem + j->dst1, j->in1 + (size_t)r * j->in1_stride, j->in1_bytes);
            j->run(mem, j->n);
            if (j->out0) memcpy(j->out0 + (size_t)r * j->out0_stride, mem + j->src0, j->out0_bytes);
            if (j->out1) memcpy(j->out1 + (size_t)r * j->out1_stride, mem + j->src1, j->out1_bytes);
        }
    }
    j->seconds = now_s() - t0;
    free(mem);
    return 0;
}

/* Runs rows [0, rows) `reps` times across `threads` pthreads; returns wall seconds (max over threads). */
double watref_run_batch(wat_fn precompute, wat_fn run, uint32_t pages, uint32_t n,
                        const void *in0, size_t in0_bytes, size_t in0_stride, uint32_t dst0,
                        const void *in1, size_t in1_bytes, size_t in1_stride, uint32_t dst1,
                        void *out0, size_t out0_bytes, size_t out0_stride, uint32_t src0,
                        void *out1, size_t out1_bytes, size_t out1_stride, uint32_t src1,
                        long rows, long reps, int threads) {
    if (threads < 1) threads = 1;
    if (threads > rows) threads = (int)rows;
    job_t *jobs = (job_t *)calloc((size_t)threads, sizeof(job_t));
    pthread_t *tid = (pthread_t *)calloc((size_t)threads, sizeof(pthread_t));
    for (int t = 0; t < threads; t++) {
        job_t *j = &jobs[t];
        j->precompute = precompute; j->run = run; j->pages = pages; j->n = n;
        j->in0 = (const uint8_t *)in0; j->in0_bytes = in0_bytes; j->in0_stride = in0_stride; j->dst0 = dst0;
        j->in1 = (const uint8_t *)in1; j->in1_bytes = in1_bytes; j->in1_stride = in1_stride; j->dst1 = dst1;
        j->out0 = (uint8_t *)out0; j->out0_bytes = out0_bytes; j->out0_stride = out0_stride; j->src0 = src0;
        j->out1 = (uint8_t *)out1; j->out1_bytes = out1_bytes; j->out1_stride = out1_stride; j->src1 = src1;
        j->row_begin = rows * t / threads; j->row_end = rows * (t + 1) / threads; j->reps = reps;
    }
    double t0 = now_s();
    for (int t = 0; t < threads; t++) pthread_create(&tid[t], 0, worker, &jobs[t]);
    for (int t = 0; t < threads; t++) pthread_join(tid[t], 0);
    double wall = now_s() - t0;
    free(jobs); free(tid);
    return wall;
}
