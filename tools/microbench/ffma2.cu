// Microbenchmark: issue/throughput of scalar FFMA vs packed FFMA2 (sm_100a) with and without
// interleaved integer work.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2 ffma2.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE> __global__ void k(float *out, int iters, float seed) {
    float2 a[8];
    int acc[4] = {threadIdx.x, threadIdx.x + 1, threadIdx.x + 2, threadIdx.x + 3};
    for (int i = 0; i < 8; i++) a[i] = make_float2(seed + i, seed - i);
    float2 m = make_float2(1.0001f, 0.9999f), c = make_float2(0.5f, 0.25f);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                if (MODE == 0 || MODE == 2) {          // scalar: 2 FFMA per complex value
                    a[i].x = fmaf(a[i].x, m.x, c.x);
                    a[i].y = fmaf(a[i].y, m.y, c.y);
                } else {                                // packed: 1 FFMA2
                    a[i] = __ffma2_rn(a[i], m, c);
                }
                if (MODE >= 2) {                        // plus one integer op per FP pair
                    acc[i & 3] = (acc[i & 3] ^ (acc[(i + 1) & 3] >> 3)) + it;
                }
            }
        }
    }
    float s = 0;
    for (int i = 0; i < 8; i++) s += a[i].x + a[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + acc[0] + acc[1] + acc[2] + acc[3];
}

template <int MODE> void run(const char *name) {
    float *out;
    cudaMalloc(&out, 148 * 8 * 1024 * sizeof(float));
    int iters = 4096;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148 * 4, 512>>>(out, 16, 1.0f);
    cudaEventRecord(e0);
    k<MODE><<<148 * 4, 512>>>(out, iters, 1.0f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double lane_fma = 148.0 * 4 * 512 * (double)iters * 4 * 8 * 2;   // scalar-equivalent FMAs
    printf("%-28s %.3f ms  %.1f TFLOP/s (fp32 fma=2flop)\n", name, ms, lane_fma * 2 / ms / 1e9);
    cudaFree(out);
}

int main() {
    run<0>("scalar FFMA");
    run<1>("packed FFMA2");
    run<2>("scalar FFMA + int");
    run<3>("packed FFMA2 + int");
    return 0;
}
