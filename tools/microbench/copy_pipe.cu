// Microbenchmark: ceiling of the TMA-pipelined streaming structure used by k_c2c_pipe, without the FFT.
//   mode 0: persistent CTAs, STAGES-deep cp.async.bulk prefetch, LDS.32 -> STG.32 (what the FFT kernel does for I/O)
//   mode 1: same, but results go back through smem and a TMA bulk store
//   mode 2: plain grid-stride float4 copy (reference point)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o copy_pipe copy_pipe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, int c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect(uint64_t *b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t ph) {
    asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}" ::"r"(s32(b)), "r"(ph) : "memory");
}
__device__ __forceinline__ void tma_ld(void *d, const void *s, uint32_t n, uint64_t *b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(d)), "l"(s), "r"(n), "r"(s32(b)) : "memory");
}
__device__ __forceinline__ void tma_st(void *d, const void *s, uint32_t n) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(d), "r"(s32(s)), "r"(n) : "memory");
}

template <int STAGES, int MODE, int TILE_FLOATS, int THREADS>
__global__ void __launch_bounds__(THREADS) k(const float *in, float *out, long tiles) {
    extern __shared__ __align__(128) unsigned char sm[];
    float *buf = reinterpret_cast<float *>(sm);
    uint64_t *mb = reinterpret_cast<uint64_t *>(sm + (size_t)(STAGES + (MODE == 1 ? 1 : 0)) * TILE_FLOATS * 4);
    if (threadIdx.x == 0) { for (int i = 0; i < STAGES; i++) mbar_init(mb + i, 1); asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
    __syncthreads();
    long t = blockIdx.x;
    if (threadIdx.x == 0)
        for (int s = 0; s < STAGES - 1; s++) {
            long tt = t + (long)s * gridDim.x;
            if (tt < tiles) { mbar_expect(mb + s, TILE_FLOATS * 4); tma_ld(buf + (size_t)s * TILE_FLOATS, in + tt * TILE_FLOATS, TILE_FLOATS * 4, mb + s); }
        }
    constexpr int PER = TILE_FLOATS / THREADS;
    for (long it = 0; t < tiles; t += gridDim.x, it++) {
        const int st = it % STAGES;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (threadIdx.x == 0) {
            long tt = t + (long)(STAGES - 1) * gridDim.x;
            int s2 = (it + STAGES - 1) % STAGES;
            if (tt < tiles) { mbar_expect(mb + s2, TILE_FLOATS * 4); tma_ld(buf + (size_t)s2 * TILE_FLOATS, in + tt * TILE_FLOATS, TILE_FLOATS * 4, mb + s2); }
        }
        mbar_wait(mb + st, (it / STAGES) & 1);
        const float *src = buf + (size_t)st * TILE_FLOATS;
        float v[PER];
#pragma unroll
        for (int e = 0; e < PER; e++) v[e] = src[threadIdx.x + e * THREADS] * 1.0001f;
        if (MODE == 0) {
            float *dst = out + t * TILE_FLOATS;
#pragma unroll
            for (int e = 0; e < PER; e++) __stcs(dst + threadIdx.x + e * THREADS, v[e]);
        } else if (MODE == 4) {
            float *dst = out + t * TILE_FLOATS;
#pragma unroll
            for (int e = 0; e < PER; e++) dst[threadIdx.x + e * THREADS] = v[e];
        } else if (MODE == 5) {
            float *dst = out + t * TILE_FLOATS;
#pragma unroll
            for (int e = 0; e < PER; e++) __stcg(dst + threadIdx.x + e * THREADS, v[e]);
        } else if (MODE == 3) {
            float4 *dst = reinterpret_cast<float4 *>(out + t * TILE_FLOATS);
            const float4 *s4 = reinterpret_cast<const float4 *>(src);
#pragma unroll
            for (int e = 0; e < PER / 4; e++) { float4 q = s4[threadIdx.x + e * THREADS]; q.x *= 1.0001f; __stcs(dst + threadIdx.x + e * THREADS, q); }
        } else {
            float *ob = buf + (size_t)STAGES * TILE_FLOATS;
            if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            __syncthreads();
#pragma unroll
            for (int e = 0; e < PER; e++) ob[threadIdx.x + e * THREADS] = v[e];
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncthreads();
            if (threadIdx.x == 0) { tma_st(out + t * TILE_FLOATS, ob, TILE_FLOATS * 4); asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
        }
    }
    if (MODE == 1 && threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

__global__ void kchunk(const float4 *in, float4 *out) {   // one 32 KB chunk per CTA, 256 threads x 8 float4
    const float4 *s = in + (long)blockIdx.x * 2048; float4 *d = out + (long)blockIdx.x * 2048;
    float4 v[8];
#pragma unroll
    for (int e = 0; e < 8; e++) v[e] = __ldcs(s + threadIdx.x + e * 256);
#pragma unroll
    for (int e = 0; e < 8; e++) { v[e].x *= 1.0001f; __stcs(d + threadIdx.x + e * 256, v[e]); }
}
// non-persistent: one CTA per 32 KB tile, TMA load -> LDS.32 -> STG.32
__global__ void __launch_bounds__(256) ktma1(const float *in, float *out) {
    extern __shared__ __align__(128) unsigned char sm[];
    float *buf = reinterpret_cast<float *>(sm);
    uint64_t *mb = reinterpret_cast<uint64_t *>(sm + 8192 * 4);
    if (threadIdx.x == 0) { mbar_init(mb, 1); asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); mbar_expect(mb, 8192 * 4); tma_ld(buf, in + (long)blockIdx.x * 8192, 8192 * 4, mb); }
    __syncthreads();
    mbar_wait(mb, 0);
    float *dst = out + (long)blockIdx.x * 8192;
#pragma unroll
    for (int e = 0; e < 32; e++) __stcs(dst + threadIdx.x + e * 256, buf[threadIdx.x + e * 256] * 1.0001f);
}
// persistent: CTAs loop over 32 KB chunks strided by the grid, LDG.128 (no TMA)
__global__ void __launch_bounds__(256) kchunk_persist(const float4 *in, float4 *out, long tiles) {
    for (long t = blockIdx.x; t < tiles; t += gridDim.x) {
        const float4 *s = in + t * 2048; float4 *d = out + t * 2048;
        float4 v[8];
#pragma unroll
        for (int e = 0; e < 8; e++) v[e] = __ldcs(s + threadIdx.x + e * 256);
#pragma unroll
        for (int e = 0; e < 8; e++) { v[e].x *= 1.0001f; __stcs(d + threadIdx.x + e * 256, v[e]); }
    }
}
// persistent + dynamic (atomic) tile claims: tiles are handed out in monotonic order like the HW CTA scheduler
__global__ void __launch_bounds__(256) kchunk_dyn(const float4 *in, float4 *out, long tiles, unsigned long long *ctr) {
    __shared__ long next;
    for (;;) {
        if (threadIdx.x == 0) next = (long)atomicAdd(ctr, 1ULL);
        __syncthreads();
        const long t = next;
        __syncthreads();
        if (t >= tiles) break;
        const float4 *s = in + t * 2048; float4 *d = out + t * 2048;
        float4 v[8];
#pragma unroll
        for (int e = 0; e < 8; e++) v[e] = __ldcs(s + threadIdx.x + e * 256);
#pragma unroll
        for (int e = 0; e < 8; e++) { v[e].x *= 1.0001f; __stcs(d + threadIdx.x + e * 256, v[e]); }
    }
}
// TMA 2-stage pipeline with dynamic tile claims (claim one tile ahead, when issuing its prefetch)
__global__ void __launch_bounds__(256) kpipe_dyn(const float *in, float *out, long tiles, unsigned long long *ctr) {
    extern __shared__ __align__(128) unsigned char sm[];
    float *buf = reinterpret_cast<float *>(sm);
    uint64_t *mb = reinterpret_cast<uint64_t *>(sm + 2 * 8192 * 4);
    long *slot = reinterpret_cast<long *>(sm + 2 * 8192 * 4 + 32);   // claimed tile per stage
    if (threadIdx.x == 0) {
        mbar_init(mb, 1); mbar_init(mb + 1, 1); asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        long t0 = (long)atomicAdd(ctr, 1ULL); slot[0] = t0;
        if (t0 < tiles) { mbar_expect(mb, 8192 * 4); tma_ld(buf, in + t0 * 8192, 8192 * 4, mb); }
    }
    __syncthreads();
    for (long it = 0;; it++) {
        const int st = it & 1;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        const long t = slot[st];
        if (t >= tiles) break;
        if (threadIdx.x == 0) {
            long tn = (long)atomicAdd(ctr, 1ULL); slot[st ^ 1] = tn;
            if (tn < tiles) { mbar_expect(mb + (st ^ 1), 8192 * 4); tma_ld(buf + (size_t)(st ^ 1) * 8192, in + tn * 8192, 8192 * 4, mb + (st ^ 1)); }
        }
        mbar_wait(mb + st, (it >> 1) & 1);
        const float *src = buf + (size_t)st * 8192;
        float *dst = out + t * 8192;
#pragma unroll
        for (int e = 0; e < 32; e++) __stcs(dst + threadIdx.x + e * 256, src[threadIdx.x + e * 256] * 1.0001f);
    }
}
__global__ void kplain(const float4 *in, float4 *out, long n) {
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        float4 v = __ldcs(in + i); v.x *= 1.0001f; __stcs(out + i, v);
    }
}

template <int STAGES, int MODE, int TILE_FLOATS, int THREADS> void run(const char *name, const float *in, float *out, long nfloat) {
    size_t smem = (size_t)(STAGES + (MODE == 1 ? 1 : 0)) * TILE_FLOATS * 4 + 64;
    auto kern = k<STAGES, MODE, TILE_FLOATS, THREADS>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int per = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, kern, THREADS, smem);
    long tiles = nfloat / TILE_FLOATS;
    int grid = per * 148;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; i++) kern<<<grid, THREADS, smem>>>(in, out, tiles);
    float best = 1e9;
    for (int i = 0; i < 10; i++) { cudaEventRecord(e0); kern<<<grid, THREADS, smem>>>(in, out, tiles); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
    printf("%-44s ctas/SM %d  %.4f ms  %.0f GB/s  (%s)\n", name, per, best, 2.0 * nfloat * 4 / best / 1e6, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    long nfloat = 1L << 28;   // 1 GiB in, 1 GiB out
    float *in, *out; cudaMalloc(&in, nfloat * 4); cudaMalloc(&out, nfloat * 4); cudaMemset(in, 0, nfloat * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; i++) kplain<<<148 * 16, 256>>>((const float4 *)in, (float4 *)out, nfloat / 4);
    float best = 1e9;
    for (int i = 0; i < 10; i++) { cudaEventRecord(e0); kplain<<<148 * 16, 256>>>((const float4 *)in, (float4 *)out, nfloat / 4); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
    printf("%-44s %.4f ms  %.0f GB/s\n", "plain float4 grid-stride copy", best, 2.0 * nfloat * 4 / best / 1e6);
    best = 1e9;
    for (int i = 0; i < 13; i++) { cudaEventRecord(e0); kchunk<<<(unsigned)(nfloat / 8192), 256>>>((const float4 *)in, (float4 *)out); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (i > 2 && ms < best) best = ms; }
    printf("%-44s %.4f ms  %.0f GB/s\n", "non-persistent 32KB chunk per CTA, float4", best, 2.0 * nfloat * 4 / best / 1e6);
    cudaFuncSetAttribute(ktma1, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 4 + 64);
    best = 1e9;
    for (int i = 0; i < 13; i++) { cudaEventRecord(e0); ktma1<<<(unsigned)(nfloat / 8192), 256, 8192 * 4 + 64>>>(in, out); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (i > 2 && ms < best) best = ms; }
    printf("%-44s %.4f ms  %.0f GB/s\n", "non-persistent, TMA load 32KB, STG.32", best, 2.0 * nfloat * 4 / best / 1e6);
    for (int g : {148 * 2, 148 * 4, 148 * 8}) {
        best = 1e9;
        for (int i = 0; i < 13; i++) { cudaEventRecord(e0); kchunk_persist<<<g, 256>>>((const float4 *)in, (float4 *)out, nfloat / 8192); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (i > 2 && ms < best) best = ms; }
        printf("persistent 32KB chunks LDG.128, grid %-5d        %.4f ms  %.0f GB/s\n", g, best, 2.0 * nfloat * 4 / best / 1e6);
    }
    unsigned long long *ctr; cudaMalloc(&ctr, 8);
    for (int g : {148 * 4, 148 * 8}) {
        best = 1e9;
        for (int i = 0; i < 13; i++) { cudaMemsetAsync(ctr, 0, 8); cudaEventRecord(e0); kchunk_dyn<<<g, 256>>>((const float4 *)in, (float4 *)out, nfloat / 8192, ctr); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (i > 2 && ms < best) best = ms; }
        printf("persistent DYNAMIC 32KB chunks LDG.128, grid %-5d %.4f ms  %.0f GB/s\n", g, best, 2.0 * nfloat * 4 / best / 1e6);
    }
    cudaFuncSetAttribute(kpipe_dyn, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 8192 * 4 + 64);
    for (int g : {148 * 2, 148 * 3}) {
        best = 1e9;
        for (int i = 0; i < 13; i++) { cudaMemsetAsync(ctr, 0, 8); cudaEventRecord(e0); kpipe_dyn<<<g, 256, 2 * 8192 * 4 + 64>>>(in, out, nfloat / 8192, ctr); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (i > 2 && ms < best) best = ms; }
        printf("pipe 2-stage TMA, DYNAMIC claims, grid %-5d       %.4f ms  %.0f GB/s  (%s)\n", g, best, 2.0 * nfloat * 4 / best / 1e6, cudaGetErrorString(cudaGetLastError()));
    }
    run<2, 0, 8192, 256>("pipe 2-stage 32KB tile, STG.32", in, out, nfloat);
    run<3, 0, 8192, 256>("pipe 3-stage 32KB tile, STG.32", in, out, nfloat);
    run<2, 0, 4096, 128>("pipe 2-stage 16KB tile/128thr, STG.32", in, out, nfloat);
    run<4, 0, 4096, 128>("pipe 4-stage 16KB tile/128thr, STG.32", in, out, nfloat);
    run<2, 0, 2048, 64>("pipe 2-stage 8KB tile/64thr, STG.32", in, out, nfloat);
    run<2, 3, 8192, 256>("pipe 2-stage 32KB tile, STG.128 .cs", in, out, nfloat);
    run<2, 4, 8192, 256>("pipe 2-stage 32KB tile, STG.32 default", in, out, nfloat);
    run<2, 5, 8192, 256>("pipe 2-stage 32KB tile, STG.32 .cg", in, out, nfloat);
    run<2, 1, 8192, 256>("pipe 2-stage 32KB tile, TMA store", in, out, nfloat);
    run<3, 1, 8192, 256>("pipe 3-stage 32KB tile, TMA store", in, out, nfloat);
    run<2, 1, 4096, 128>("pipe 2-stage 16KB tile/128thr, TMA store", in, out, nfloat);
    return 0;
}
