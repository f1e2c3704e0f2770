// pcie_peak.cu -- concurrent pinned-copy ceiling of the host links of one box.
// One process, one host thread per GPU.  For G = 1, 2, 4, 8 GPUs at once and for each way of obtaining the pinned host
// memory, times plain cudaMemcpyAsync H2D alone, D2H alone and both at once (one stream each), CUDA-event timed per
// GPU, and prints one JSON line per case with the per-GPU minimum and the sum over GPUs.
//   alloc modes: hostalloc        cudaHostAlloc(Default), from the worker thread of that GPU
//                hostalloc_wc     cudaHostAllocWriteCombined (H2D source only; no snooping of CPU caches)
//                register_thp     aligned_alloc + madvise(MADV_HUGEPAGE) + first touch by the worker + cudaHostRegister
//   affinity:    none | spread    (worker i pinned to the i-th slice of the allowed CPUs before it allocates)
//   device order: seq     GPUs 0 .. G-1
//                 spread  GPUs i * (all / G): on a two-root-complex host the first half of the GPUs hangs off one
//                         complex and the second half off the other, so G < all GPUs in `seq` order share ONE complex
// Build: make -C tools/microbench pcie_peak
// Run:   tools/microbench/pcie_peak [MiB per copy = 256] [copies = 6] [orders = seq | spread | both | pairs] [alloc mode | all]
#include <cuda_runtime.h>
#include <pthread.h>
#include <sched.h>
#include <sys/mman.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

struct Result { double h2d, d2h, dup_h2d, dup_d2h; };
static pthread_barrier_t g_bar;

static void pin_slice(int i, int n) {
    cpu_set_t allowed;
    if (sched_getaffinity(0, sizeof allowed, &allowed) != 0) return;
    std::vector<int> cpus;
    for (int c = 0; c < CPU_SETSIZE; c++) if (CPU_ISSET(c, &allowed)) cpus.push_back(c);
    const int per = std::max<int>(1, (int)cpus.size() / n);
    cpu_set_t mine; CPU_ZERO(&mine);
    for (int k = i * per; k < (i + 1) * per && k < (int)cpus.size(); k++) CPU_SET(cpus[k], &mine);
    pthread_setaffinity_np(pthread_self(), sizeof mine, &mine);
}

static void worker(int slot, int dev, int ngpu, const std::string &mode, bool spread, size_t bytes, int copies, Result *out) {
    if (spread) pin_slice(slot, ngpu);
    CK(cudaSetDevice(dev));
    void *h[2], *d[2];
    for (int i = 0; i < 2; i++) {
        if (mode == "hostalloc") CK(cudaHostAlloc(&h[i], bytes, cudaHostAllocDefault));
        else if (mode == "hostalloc_wc") CK(cudaHostAlloc(&h[i], bytes, i == 0 ? cudaHostAllocWriteCombined : cudaHostAllocDefault));
        else {
            h[i] = aligned_alloc(2u << 20, bytes);
            madvise(h[i], bytes, MADV_HUGEPAGE);
            memset(h[i], 1, bytes);
            CK(cudaHostRegister(h[i], bytes, cudaHostRegisterDefault));
        }
        if (mode != "register_thp") memset(h[i], 1, bytes);
        CK(cudaMalloc(&d[i], bytes));
    }
    cudaStream_t s[2]; cudaEvent_t ev[4];
    for (int i = 0; i < 2; i++) CK(cudaStreamCreateWithFlags(&s[i], cudaStreamNonBlocking));
    for (int i = 0; i < 4; i++) CK(cudaEventCreate(&ev[i]));
    auto run = [&](bool up, bool down, double *tu, double *td) {
        for (int w = 0; w < 2; w++) {
            if (w == 1) pthread_barrier_wait(&g_bar);          // all GPUs start the timed pass together
            if (up) CK(cudaEventRecord(ev[0], s[0]));
            if (down) CK(cudaEventRecord(ev[2], s[1]));
            for (int i = 0; i < (w ? copies : 1); i++) {
                if (up) CK(cudaMemcpyAsync(d[0], h[0], bytes, cudaMemcpyHostToDevice, s[0]));
                if (down) CK(cudaMemcpyAsync(h[1], d[1], bytes, cudaMemcpyDeviceToHost, s[1]));
            }
            if (up) CK(cudaEventRecord(ev[1], s[0]));
            if (down) CK(cudaEventRecord(ev[3], s[1]));
            CK(cudaStreamSynchronize(s[0])); CK(cudaStreamSynchronize(s[1]));
        }
        float ms;
        if (up) { CK(cudaEventElapsedTime(&ms, ev[0], ev[1])); *tu = ms * 1e-3; }
        if (down) { CK(cudaEventElapsedTime(&ms, ev[2], ev[3])); *td = ms * 1e-3; }
    };
    double tu = 1, td = 1, du = 1, dd = 1, x;
    run(true, false, &tu, &x); run(false, true, &x, &td); run(true, true, &du, &dd);
    const double gb = (double)bytes * copies / 1e9;
    *out = Result{gb / tu, gb / td, gb / du, gb / dd};
    for (int i = 0; i < 2; i++) {
        if (mode == "register_thp") { cudaHostUnregister(h[i]); free(h[i]); } else cudaFreeHost(h[i]);
        cudaFree(d[i]); cudaStreamDestroy(s[i]);
    }
    for (int i = 0; i < 4; i++) cudaEventDestroy(ev[i]);
}

int main(int argc, char **argv) {
    const size_t bytes = (size_t)(argc > 1 ? atoi(argv[1]) : 256) << 20;
    const int copies = argc > 2 ? atoi(argv[2]) : 6;
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    for (int d = 0; d < ndev; d++) { CK(cudaSetDevice(d)); CK(cudaFree(0)); }
    const std::string orders = argc > 3 ? argv[3] : "seq";
    const std::string only = argc > 4 ? argv[4] : "all";
    const char *modes[] = {"hostalloc", "hostalloc_wc", "register_thp"};
    // device sets: seq / spread for G = 1, 2, 4, ...; `pairs` = {0, k} for every k (which GPUs share a host bridge with GPU 0?)
    std::vector<std::vector<int>> sets;
    if (orders == "pairs") {
        for (int k = 1; k < ndev; k++) sets.push_back({0, k});
    } else {
        for (int order = 0; order < 2; order++)
            for (int g = 1; g <= ndev; g *= 2) {
                if ((order == 0 && orders == "spread") || (order == 1 && (orders == "seq" || g == 1 || g == ndev))) continue;
                std::vector<int> v;
                for (int d = 0; d < g; d++) v.push_back(order ? d * (ndev / g) : d);
                sets.push_back(v);
            }
    }
    for (const std::vector<int> &set : sets)
        for (const char *mode : modes)
            for (int spread = 0; spread < 2; spread++) {
                const int g = (int)set.size();
                if (spread && g == 1) continue;
                if (only != "all" && only != mode) continue;
                pthread_barrier_init(&g_bar, nullptr, g);
                std::vector<Result> r(g);
                std::vector<std::thread> th;
                std::string devs;
                for (int d = 0; d < g; d++) {
                    devs += (d ? "," : "") + std::to_string(set[d]);
                    th.emplace_back(worker, d, set[d], g, std::string(mode), spread != 0, bytes, copies, &r[d]);
                }
                for (auto &t : th) t.join();
                pthread_barrier_destroy(&g_bar);
                auto agg = [&](double Result::*f, double *mn, double *sum) { *mn = 1e30; *sum = 0; for (auto &x : r) { *mn = std::min(*mn, x.*f); *sum += x.*f; } };
                double a[8];
                agg(&Result::h2d, &a[0], &a[1]); agg(&Result::d2h, &a[2], &a[3]); agg(&Result::dup_h2d, &a[4], &a[5]); agg(&Result::dup_d2h, &a[6], &a[7]);
                printf("{\"gpus\": %d, \"devices\": \"%s\", \"alloc\": \"%s\", \"affinity\": \"%s\", \"MiB\": %zu, \"h2d_min\": %.2f, \"h2d_sum\": %.2f, \"d2h_min\": %.2f, \"d2h_sum\": %.2f, "
                       "\"duplex_h2d_min\": %.2f, \"duplex_h2d_sum\": %.2f, \"duplex_d2h_min\": %.2f, \"duplex_d2h_sum\": %.2f}\n",
                       g, devs.c_str(), mode, spread ? "spread" : "none", bytes >> 20, a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7]);
                fflush(stdout);
            }
    return 0;
}
