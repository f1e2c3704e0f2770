/*
 * watref_threads.c -- pthread driver around the transpiled reference modules
 * (TEST INFRASTRUCTURE; links into oracle/_ref/libwatref.so).
 *
 * The reference is single-threaded: one WebAssembly.Instance with private
 * linear memory per context (index.js:13-18).  The CPU baseline named by
 * BASELINE.json ("reference WASM on all host cores via worker_threads") is
 * restated here as a PERSISTENT pool of pthreads, one private module memory
 * per thread (like one instance per worker).  As in the reference's harness
 * (benchmarks/lib/wat-contexts.js:110-131) the instance is created and its
 * precompute export is called ONCE per size, outside the timed region
 * (watref_pool_prepare); the timed region (watref_pool_run) is only
 *     memcpy(input row) -> transform(n)
 * per row, exactly what bench() charges (benchmarks/lib/wat-contexts.js:125-129).
 * Wall time is taken from the moment the workers are released to the moment
 * the last one reports back; no allocation, memset or table building inside.
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <sched.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <math.h>

typedef void (*wat_fn)(uint8_t *, uint32_t);

typedef struct {
    wat_fn fn;                 /* precompute (prepare) or transform (run) */
    int prepare;               /* 1: (re)allocate the module memory and call fn once */
    uint32_t pages, n;
    const uint8_t *in0; size_t in0_bytes, in0_stride; uint32_t dst0;   /* plane 0 -> memory offset dst0 */
    const uint8_t *in1; size_t in1_bytes, in1_stride; uint32_t dst1;   /* optional plane 1 */
    uint8_t *out0; size_t out0_bytes, out0_stride; uint32_t src0;      /* optional copy-out */
    uint8_t *out1; size_t out1_bytes, out1_stride; uint32_t src1;
    long rows, reps;
    /* spectrogram job (stft != 0): rows = frames */
    int stft, hop, wsize;
    const float *samples; const double *window; float *spec;
    double db_floor, range;
} job_t;

typedef struct pool pool_t;
typedef struct {
    pool_t *pool;
    int id;
    pthread_t th;
    uint8_t *mem;
    size_t mem_bytes;
} worker_t;

struct pool {
    int threads;
    worker_t *w;
    pthread_mutex_t mu;
    pthread_cond_t go, done;
    long generation;
    int pending, quit;
    job_t job;
};

static double now_s(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }

static void do_job(worker_t *w, const job_t *j, int threads) {
    if (j->prepare) {
        const size_t bytes = (size_t)j->pages * 65536;
        if (w->mem_bytes != bytes) {
            free(w->mem);
            w->mem = (uint8_t *)aligned_alloc(64, bytes);
            w->mem_bytes = bytes;
        }
        memset(w->mem, 0, bytes);
        j->fn(w->mem, j->n);
        return;
    }
    uint8_t *mem = w->mem;
    const long r0 = j->rows * w->id / threads, r1 = j->rows * (w->id + 1) / threads;
    if (j->stft) {
        /* the per-frame loop of generateSpectrogram (playground/src/spectrogram.js:299-353) around the module's rfft:
         * slice -> applyWindow (:37-47, product of doubles rounded to f32; the window is tabulated here instead of
         * calling Math.cos per sample, which only helps the CPU) -> zeroPad (:54-59) -> inputBuffer.set -> run ->
         * computeMagnitude (:64-75, rounded to f32) -> /(fftSize/2) -> 20 log10(m + 1e-10) -> gain/range clamp,
         * bins 0..2 zeroed (:336-352) */
        float *in = (float *)mem;
        const int n = (int)j->n, bins = n / 2 + 1;
        const double half = n / 2;
        for (long rep = 0; rep < j->reps; rep++)
            for (long f = r0; f < r1; f++) {
                const float *src = j->samples + f * (long)j->hop;
                for (int i = 0; i < j->wsize; i++) in[i] = (float)((double)src[i] * j->window[i]);
                for (int i = j->wsize; i < n; i++) in[i] = 0.0f;
                j->fn(mem, j->n);
                float *o = j->spec + f * (long)bins;
                for (int b = 0; b < bins; b++) {
                    if (b < 3) { o[b] = 0.0f; continue; }
                    const double re = in[2 * b], im = in[2 * b + 1];
                    const float mag = (float)sqrt(re * re + im * im);
                    const double db = 20.0 * log10((double)mag / half + 1e-10);
                    double v = (db - j->db_floor) / j->range;
                    v = v < 0 ? 0 : (v > 1 ? 1 : v);
                    o[b] = (float)v;
                }
            }
        return;
    }
    for (long rep = 0; rep < j->reps; rep++) {
        for (long r = r0; r < r1; r++) {
            memcpy(mem + j->dst0, j->in0 + (size_t)r * j->in0_stride, j->in0_bytes);
            if (j->in1) memcpy(mem + j->dst1, j->in1 + (size_t)r * j->in1_stride, j->in1_bytes);
            j->fn(mem, j->n);
            if (j->out0) memcpy(j->out0 + (size_t)r * j->out0_stride, mem + j->src0, j->out0_bytes);
            if (j->out1) memcpy(j->out1 + (size_t)r * j->out1_stride, mem + j->src1, j->out1_bytes);
        }
    }
}

static void *worker_main(void *arg) {
    worker_t *w = (worker_t *)arg;
    pool_t *p = w->pool;
    long seen = 0;
    for (;;) {
        pthread_mutex_lock(&p->mu);
        while (p->generation == seen && !p->quit) pthread_cond_wait(&p->go, &p->mu);
        if (p->quit) { pthread_mutex_unlock(&p->mu); break; }
        seen = p->generation;
        const job_t job = p->job;
        pthread_mutex_unlock(&p->mu);
        do_job(w, &job, p->threads);
        pthread_mutex_lock(&p->mu);
        if (--p->pending == 0) pthread_cond_signal(&p->done);
        pthread_mutex_unlock(&p->mu);
    }
    return 0;
}

/* releases the workers on `job` and waits for all of them; returns the wall seconds in between */
static double dispatch(pool_t *p, const job_t *job) {
    pthread_mutex_lock(&p->mu);
    p->job = *job;
    p->pending = p->threads;
    p->generation++;
    const double t0 = now_s();
    pthread_cond_broadcast(&p->go);
    while (p->pending) pthread_cond_wait(&p->done, &p->mu);
    const double t1 = now_s();
    pthread_mutex_unlock(&p->mu);
    return t1 - t0;
}

void *watref_pool_create(int threads) {
    if (threads < 1) threads = 1;
    pool_t *p = (pool_t *)calloc(1, sizeof(pool_t));
    p->threads = threads;
    p->w = (worker_t *)calloc((size_t)threads, sizeof(worker_t));
    pthread_mutex_init(&p->mu, 0);
    pthread_cond_init(&p->go, 0);
    pthread_cond_init(&p->done, 0);
    cpu_set_t allowed;
    int ncpu = 0, cpus[1024];
    if (sched_getaffinity(0, sizeof allowed, &allowed) == 0)
        for (int c = 0; c < CPU_SETSIZE && ncpu < 1024; c++) if (CPU_ISSET(c, &allowed)) cpus[ncpu++] = c;
    for (int t = 0; t < threads; t++) {
        p->w[t].pool = p; p->w[t].id = t;
        pthread_create(&p->w[t].th, 0, worker_main, &p->w[t]);
        if (ncpu >= threads) {          /* one worker per allowed core: no migration, private L1/L2 like one worker_thread each */
            cpu_set_t one; CPU_ZERO(&one); CPU_SET(cpus[t], &one);
            pthread_setaffinity_np(p->w[t].th, sizeof one, &one);
        }
    }
    return p;
}

void watref_pool_destroy(void *pool) {
    pool_t *p = (pool_t *)pool;
    if (!p) return;
    pthread_mutex_lock(&p->mu);
    p->quit = 1;
    pthread_cond_broadcast(&p->go);
    pthread_mutex_unlock(&p->mu);
    for (int t = 0; t < p->threads; t++) { pthread_join(p->w[t].th, 0); free(p->w[t].mem); }
    free(p->w);
    free(p);
}

int watref_pool_threads(void *pool) { return pool ? ((pool_t *)pool)->threads : 0; }

/* Untimed: every worker gets a zeroed private module memory of `pages` WASM pages and calls precompute(n) on it. */
void watref_pool_prepare(void *pool, wat_fn precompute, uint32_t pages, uint32_t n) {
    job_t j; memset(&j, 0, sizeof j);
    j.fn = precompute; j.prepare = 1; j.pages = pages; j.n = n;
    dispatch((pool_t *)pool, &j);
}

/* Timed: rows [0, rows) split contiguously over the workers, the whole set `reps` times; returns wall seconds. */
double watref_pool_run(void *pool, wat_fn run, uint32_t n,
                       const void *in0, size_t in0_bytes, size_t in0_stride, uint32_t dst0,
                       const void *in1, size_t in1_bytes, size_t in1_stride, uint32_t dst1,
                       void *out0, size_t out0_bytes, size_t out0_stride, uint32_t src0,
                       void *out1, size_t out1_bytes, size_t out1_stride, uint32_t src1,
                       long rows, long reps) {
    job_t j; memset(&j, 0, sizeof j);
    j.fn = run; j.n = n;
    j.in0 = (const uint8_t *)in0; j.in0_bytes = in0_bytes; j.in0_stride = in0_stride; j.dst0 = dst0;
    j.in1 = (const uint8_t *)in1; j.in1_bytes = in1_bytes; j.in1_stride = in1_stride; j.dst1 = dst1;
    j.out0 = (uint8_t *)out0; j.out0_bytes = out0_bytes; j.out0_stride = out0_stride; j.src0 = src0;
    j.out1 = (uint8_t *)out1; j.out1_bytes = out1_bytes; j.out1_stride = out1_stride; j.src1 = src1;
    j.rows = rows; j.reps = reps;
    return dispatch((pool_t *)pool, &j);
}

/* Timed: the spectrogram loop over `frames` frames (see do_job); `run` is the module's real transform of size n. */
double watref_pool_run_stft(void *pool, wat_fn run, uint32_t n, const float *samples, const double *window, int wsize,
                            int hop, long frames, double gain, double range, float *spec, long reps) {
    job_t j; memset(&j, 0, sizeof j);
    j.fn = run; j.n = n; j.stft = 1; j.hop = hop; j.wsize = wsize; j.samples = samples; j.window = window; j.spec = spec;
    j.db_floor = gain - range; j.range = range; j.rows = frames; j.reps = reps;
    return dispatch((pool_t *)pool, &j);
}

/* One-shot form (a transient pool; prepare is outside the returned time). */
double watref_run_batch(wat_fn precompute, wat_fn run, uint32_t pages, uint32_t n,
                        const void *in0, size_t in0_bytes, size_t in0_stride, uint32_t dst0,
                        const void *in1, size_t in1_bytes, size_t in1_stride, uint32_t dst1,
                        void *out0, size_t out0_bytes, size_t out0_stride, uint32_t src0,
                        void *out1, size_t out1_bytes, size_t out1_stride, uint32_t src1,
                        long rows, long reps, int threads) {
    if (threads < 1) threads = 1;
    if (threads > rows) threads = (int)rows;
    void *p = watref_pool_create(threads);
    watref_pool_prepare(p, precompute, pages, n);
    const double s = watref_pool_run(p, run, n, in0, in0_bytes, in0_stride, dst0, in1, in1_bytes, in1_stride, dst1,
                                     out0, out0_bytes, out0_stride, src0, out1, out1_bytes, out1_stride, src1, rows, reps);
    watref_pool_destroy(p);
    return s;
}
