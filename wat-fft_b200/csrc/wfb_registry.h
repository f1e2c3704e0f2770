// wfb_registry.h -- kernel-variant registry shared by the variant translation units and the API.
//
// Each wfb_variants_*.cu instantiates a family of kernels and returns descriptors (Variant); the
// API (wfb_api.cu) picks, per (precision, core size, kind), the variant with the highest priority
// as the plan's default and keeps the others selectable (wfb_plan_set_variant) for tuning.
#pragma once
#include "wfb_kernels.cuh"

#include <vector>

namespace wfb {

typedef cudaError_t (*launch_fn)(int io, int dir, const KParams &p, long batch, cudaStream_t s);

struct Variant {
    const char *name;
    int core_n;        // complex points of the core transform (n for C2C, n/2 for R2C)
    int threads;       // per CTA
    int rows_per_cta;  // thread groups per CTA
    size_t smem;       // dynamic shared memory per CTA
    int lanes;         // batch rows per thread group: 2 for the packed-FP32 (f32x2) kernels
    int priority;      // larger = preferred default (set from measurements, profiles/)
    int priority_inv;  // same for the inverse direction (c2r / ifft); -1 = same as `priority`
    int priority_il;   // same for interleaved c2c layouts; -1 = same as `priority`
    int align;         // required alignment (bytes) of every plane pointer: 16 for the TMA / float4 kernels,
                       // one complex value for the direct kernels (one scalar when the layout is split)
    std::vector<int> radices;
    launch_fn c2c, r2c, c2r;
    bool direct = false;   // plain global loads/stores, one launch-sized grid (no TMA pipeline, no tile counter): the zero-copy path's kernel
};

// implemented in wfb_api.cu (one launch counter / attribute cache for the whole library)
cudaError_t launch_grid(const void *kernel, size_t smem, int threads, long ctas, const KParams &p, cudaStream_t s);
cudaError_t launch_persistent(const void *kernel, size_t smem, int threads, long work_items, const KParams &p, cudaStream_t s);

const std::vector<Variant> &variants_f32_direct();
const std::vector<Variant> &variants_f32_x2();
const std::vector<Variant> &variants_f32_tile();
const std::vector<Variant> &variants_f32_pipe();
const std::vector<Variant> &variants_f64();
const std::vector<Variant> &variants_f32_real_pipe();
const std::vector<Variant> &variants_f64_pipe();

template <class PL> std::vector<int> plan_radices() {
    std::vector<int> r;
    for (int p = 0; p < PL::npass(); p++)
        for (int q = 0; q < pass_nsub(PL::code(p)); q++) r.push_back(pass_radix(PL::code(p), q));
    return r;
}

constexpr int PADQ = 16;   // one pad slot per 16 complex values: conflict-free for every plan (tools/bank_sim.py)

// one transform (or lane pair) per thread group, X groups per CTA, direct global loads/stores
template <typename R, class PL, int X, int MINB, bool SPLIT_IO> struct Launchers {
    static constexpr size_t smem = sizeof(cx<R>) * (size_t)padded_size<PADQ>(PL::N) * X;
    static constexpr int LANES = RT<R>::LANES;
    static long ctas(long batch) { return ((batch + LANES - 1) / LANES + X - 1) / X; }
    static cudaError_t c2c(int io, int dir, const KParams &p, long batch, cudaStream_t s) {
        const void *k;
        if (SPLIT_IO && io == IO_SPLIT) {
            if constexpr (SPLIT_IO)
                k = dir ? (const void *)k_c2c<R, PL, X, PADQ, IO_SPLIT, true, MINB> : (const void *)k_c2c<R, PL, X, PADQ, IO_SPLIT, false, MINB>;
            else
                k = nullptr;
        } else {
            k = dir ? (const void *)k_c2c<R, PL, X, PADQ, IO_INTERLEAVED, true, MINB> : (const void *)k_c2c<R, PL, X, PADQ, IO_INTERLEAVED, false, MINB>;
        }
        return launch_grid(k, smem, PL::T * X, ctas(batch), p, s);
    }
    static cudaError_t r2c(int, int, const KParams &p, long batch, cudaStream_t s) {
        return launch_grid((const void *)k_r2c<R, PL, X, PADQ, MINB>, smem, PL::T * X, ctas(batch), p, s);
    }
    static cudaError_t c2r(int, int, const KParams &p, long batch, cudaStream_t s) {
        return launch_grid((const void *)k_c2r<R, PL, X, PADQ, MINB>, smem, PL::T * X, ctas(batch), p, s);
    }
    static Variant make(const char *name, int priority, int priority_inv = -1, int priority_il = -1) {
        return Variant{name, PL::N, PL::T * X, X, smem, LANES, priority, priority_inv < 0 ? priority : priority_inv, priority_il < 0 ? priority : priority_il, 2 * (int)sizeof(typename RT<R>::scalar), plan_radices<PL>(), &c2c, &r2c, &c2r, true};
    }
};

// persistent TMA-pipelined c2c: grid = resident CTAs, each loops over tiles of X*LANES rows
template <typename R, class PL, int X, int MINB, int RC = 0, int PQ = PADQ, bool TS = false, bool HT = false> struct PipeLaunchers {
    static constexpr size_t smem = 2 * pipe_buf_bytes<R, PL, PQ, X>() + 64;
    static constexpr int LANES = RT<R>::LANES;
    static long tiles(long batch) { return (batch + X * LANES - 1) / (X * LANES); }
    static cudaError_t c2c(int io, int dir, const KParams &p, long batch, cudaStream_t s) {
        const void *k;
        if (io == IO_SPLIT)
            k = dir ? (const void *)k_c2c_pipe<R, PL, X, PQ, IO_SPLIT, true, MINB, RC, TS, HT> : (const void *)k_c2c_pipe<R, PL, X, PQ, IO_SPLIT, false, MINB, RC, TS, HT>;
        else
            k = dir ? (const void *)k_c2c_pipe<R, PL, X, PQ, IO_INTERLEAVED, true, MINB, RC, TS, HT> : (const void *)k_c2c_pipe<R, PL, X, PQ, IO_INTERLEAVED, false, MINB, RC, TS, HT>;
        return launch_persistent(k, smem, PL::T * X, tiles(batch), p, s);
    }
    static Variant make(const char *name, int priority, int priority_inv = -1, int priority_il = -1) {
        return Variant{name, PL::N, PL::T * X, X, smem, LANES, priority, priority_inv < 0 ? priority : priority_inv, priority_il < 0 ? priority : priority_il, 16, plan_radices<PL>(), &c2c, nullptr, nullptr};
    }
};

// register-prefetch persistent c2c (k_c2c_rpf): one transform per CTA iteration
template <typename R, class PL, int MINB, bool TS, int PQ = PADQ> struct RegPipeLaunchers {
    static constexpr size_t smem = (TS ? 2 : 1) * pipe_buf_bytes<R, PL, PQ, 1>() + 64;
    static cudaError_t c2c(int io, int dir, const KParams &p, long batch, cudaStream_t s) {
        const void *k;
        if (io == IO_SPLIT)
            k = dir ? (const void *)k_c2c_rpf<R, PL, PQ, IO_SPLIT, true, MINB, TS> : (const void *)k_c2c_rpf<R, PL, PQ, IO_SPLIT, false, MINB, TS>;
        else
            k = dir ? (const void *)k_c2c_rpf<R, PL, PQ, IO_INTERLEAVED, true, MINB, TS> : (const void *)k_c2c_rpf<R, PL, PQ, IO_INTERLEAVED, false, MINB, TS>;
        return launch_persistent(k, smem, PL::T, batch, p, s);
    }
    static Variant make(const char *name, int priority, int priority_inv = -1, int priority_il = -1) {
        // TS: bulk stores need 16-byte aligned rows; the loads are plain element loads either way
        return Variant{name, PL::N, PL::T, 1, smem, 1, priority, priority_inv < 0 ? priority : priority_inv, priority_il < 0 ? priority : priority_il, TS ? 16 : 2 * (int)sizeof(typename RT<R>::scalar), plan_radices<PL>(), &c2c, nullptr, nullptr};
    }
};

// persistent TMA-pipelined r2c / c2r (scalar lanes)
// XI: rows per tile of the c2r direction (1 = single-row tiles with shifted bulk copies, see k_real_pipe)
template <typename R, class PL, int X, int MINB, bool RC = false, int PQ = PADQ, bool TS = false, int XI = X, bool HT = false> struct RealPipeLaunchers {
    static constexpr size_t smem_f = 2 * real_pipe_buf_bytes<R, PL, PQ, X, false>() + 64;
    static constexpr size_t smem_i = 2 * real_pipe_buf_bytes<R, PL, PQ, XI, true>() + 64;
    static cudaError_t r2c(int, int, const KParams &p, long batch, cudaStream_t s) {
        return launch_persistent((const void *)k_real_pipe<R, PL, X, PQ, false, MINB, RC, TS, HT>, smem_f, PL::T * X, (batch + X - 1) / X, p, s);
    }
    static cudaError_t c2r(int, int, const KParams &p, long batch, cudaStream_t s) {
        return launch_persistent((const void *)k_real_pipe<R, PL, XI, PQ, true, MINB, false, TS, HT>, smem_i, PL::T * XI, (batch + XI - 1) / XI, p, s);
    }
    static Variant make(const char *name, int priority, int priority_inv = -1, int priority_il = -1) {
        return Variant{name, PL::N, PL::T * X, X, smem_i > smem_f ? smem_i : smem_f, 1, priority, priority_inv < 0 ? priority : priority_inv, priority_il < 0 ? priority : priority_il, 16, plan_radices<PL>(), nullptr, &r2c, &c2r};
    }
};

// thread-per-row tile kernels (N <= 64)
template <class PL, int X, int MINB> struct TileLaunchers {
    static constexpr size_t smem = sizeof(float) * 2 * (size_t)(PL::N + 2) * X;
    static cudaError_t c2c(int io, int dir, const KParams &p, long batch, cudaStream_t s) {
        const void *k;
        if (io == IO_SPLIT)
            k = dir ? (const void *)k_c2c_tile<float, PL, X, IO_SPLIT, true, MINB> : (const void *)k_c2c_tile<float, PL, X, IO_SPLIT, false, MINB>;
        else
            k = dir ? (const void *)k_c2c_tile<float, PL, X, IO_INTERLEAVED, true, MINB> : (const void *)k_c2c_tile<float, PL, X, IO_INTERLEAVED, false, MINB>;
        return launch_grid(k, smem, X, (batch + X - 1) / X, p, s);
    }
    static Variant make(const char *name, int priority, int priority_inv = -1, int priority_il = -1) {
        return Variant{name, PL::N, X, X, smem, 1, priority, priority_inv < 0 ? priority : priority_inv, priority_il < 0 ? priority : priority_il, 16, plan_radices<PL>(), &c2c, nullptr, nullptr};
    }
};

// persistent, fully TMA-fed thread-per-row c2c (N <= 64)
template <class PL, int X, int MINB, typename R = float> struct TilePipeLaunchers {
    static constexpr bool F64 = sizeof(R) == 8;
    static constexpr size_t smem_i = 2 * tpipe_buf_bytes<R, PL, X, IO_INTERLEAVED>() + 64;
    static constexpr size_t smem_s = F64 ? 0 : 2 * tpipe_buf_bytes<float, PL, X, IO_SPLIT>() + 64;
    static constexpr size_t smem = smem_s > smem_i ? smem_s : smem_i;
    static cudaError_t c2c(int io, int dir, const KParams &p, long batch, cudaStream_t s) {
        const void *k;
        if (io == IO_SPLIT) {
            if constexpr (F64) return cudaErrorInvalidDeviceFunction;      // the f64 modules are interleaved only
            else k = dir ? (const void *)k_c2c_tpipe<float, PL, X, IO_SPLIT, true, MINB> : (const void *)k_c2c_tpipe<float, PL, X, IO_SPLIT, false, MINB>;
        } else {
            k = dir ? (const void *)k_c2c_tpipe<R, PL, X, IO_INTERLEAVED, true, MINB> : (const void *)k_c2c_tpipe<R, PL, X, IO_INTERLEAVED, false, MINB>;
        }
        return launch_persistent(k, smem, X, (batch + X - 1) / X, p, s);
    }
    static Variant make(const char *name, int priority, int priority_inv = -1, int priority_il = -1) {
        return Variant{name, PL::N, X, X, smem, 1, priority, priority_inv < 0 ? priority : priority_inv, priority_il < 0 ? priority : priority_il, 16, plan_radices<PL>(), &c2c, nullptr, nullptr};
    }
};

// persistent, fully TMA-fed thread-per-row r2c / c2r (N = 64, 128)
template <class PL, int X, int MINB, typename R = float> struct RealTilePipeLaunchers {
    static constexpr size_t smem = 2 * rtpipe_buf_bytes<R, PL, X>() + 64;
    static cudaError_t r2c(int, int, const KParams &p, long batch, cudaStream_t s) {
        return launch_persistent((const void *)k_real_tpipe<R, PL, X, false, MINB>, smem, X, (batch + X - 1) / X, p, s);
    }
    static cudaError_t c2r(int, int, const KParams &p, long batch, cudaStream_t s) {
        return launch_persistent((const void *)k_real_tpipe<R, PL, X, true, MINB>, smem, X, (batch + X - 1) / X, p, s);
    }
    static Variant make(const char *name, int priority, int priority_inv = -1, int priority_il = -1) {
        return Variant{name, PL::N, X, X, smem, 1, priority, priority_inv < 0 ? priority : priority_inv, priority_il < 0 ? priority : priority_il, 16, plan_radices<PL>(), nullptr, &r2c, &c2r};
    }
};

// thread-per-row tile kernels for the real transforms (core M <= 64)
template <class PL, int X, int MINB> struct RealTileLaunchers {
    static constexpr size_t smem = sizeof(float) * (size_t)(2 * PL::N + 4) * X;
    static cudaError_t r2c(int, int, const KParams &p, long batch, cudaStream_t s) {
        return launch_grid((const void *)k_r2c_tile<PL, X, MINB>, smem, X, (batch + X - 1) / X, p, s);
    }
    static cudaError_t c2r(int, int, const KParams &p, long batch, cudaStream_t s) {
        return launch_grid((const void *)k_c2r_tile<PL, X, MINB>, smem, X, (batch + X - 1) / X, p, s);
    }
    static Variant make(const char *name, int priority, int priority_inv = -1, int priority_il = -1) {
        return Variant{name, PL::N, X, X, smem, 1, priority, priority_inv < 0 ? priority : priority_inv, priority_il < 0 ? priority : priority_il, 16, plan_radices<PL>(), nullptr, &r2c, &c2r};
    }
};

// fused STFT (frame gather + window + r2c + magnitude/dB)
typedef cudaError_t (*stft_launch_fn)(const StftParams &sp, cudaStream_t s);
struct StftVariant {
    const char *name; int core_n; std::vector<int> radices; stft_launch_fn launch, launch_pipe, launch_span;
    std::vector<int> span_radices;     // stage structure of the span kernel's core plan (its own twiddle table when it differs)
    bool span_default = false;         // measured faster than the direct kernel at hop = N/4 (profiles/): used unless WFB_STFT_SPAN says otherwise
};
// persistent launch whose dynamic shared memory varies from call to call (the span stages depend on hop and window)
cudaError_t launch_persistent_dyn(const void *kernel, size_t smem, int threads, long work_items, const StftParams &sp, cudaStream_t s);
cudaError_t launch_persistent_raw(const void *kernel, size_t smem, int threads, long work_items, void *params, cudaStream_t s);
cudaError_t launch_grid_raw(const void *kernel, size_t smem, int threads, long ctas, void *params, cudaStream_t s);
const std::vector<StftVariant> &variants_stft();
template <class PL, int X, int MINB, int PQ = PADQ> struct StftLaunchers {
    static constexpr size_t smem = sizeof(cx<float>) * (size_t)padded_size<PQ>(PL::N) * X;
    // kernel flavour: fast dB (floor above -150 dB) / exact dB / complex output; the unpadded fast-dB case (window =
    // FFT size, the spectrogram default) gets its own instantiation without the per-element bound check
    static int flavour(const StftParams &sp) {
        if (sp.mode == STFT_MODE_COMPLEX) return STFT_K_COMPLEX;
        return sp.db_floor >= -150.0f ? STFT_K_FAST : STFT_K_EXACT;
    }
    static cudaError_t launch(const StftParams &sp, cudaStream_t s) {
        const bool pad = sp.wsize < 2 * PL::N;
        const void *k;
        switch (flavour(sp)) {
            case STFT_K_FAST: k = pad ? (const void *)k_stft<PL, X, PQ, MINB, STFT_K_FAST, true> : (const void *)k_stft<PL, X, PQ, MINB, STFT_K_FAST, false>; break;
            case STFT_K_EXACT: k = (const void *)k_stft<PL, X, PQ, MINB, STFT_K_EXACT, true>; break;
            default: k = (const void *)k_stft<PL, X, PQ, MINB, STFT_K_COMPLEX, true>; break;
        }
        return launch_grid_raw(k, smem, PL::T * X, (sp.frames + X - 1) / X, (void *)&sp, s);
    }
    // pipelined variant: XP frames per CTA (>= one warp of threads)
    static constexpr int XP = (PL::T * X >= 128) ? X : 128 / PL::T;
    static constexpr size_t smem_pipe = 2 * real_pipe_buf_bytes<float, PL, PQ, XP, false>() + 64;
    static cudaError_t launch_pipe(const StftParams &sp, cudaStream_t s) {
        const bool pad = sp.wsize < 2 * PL::N;
        const void *k;
        switch (flavour(sp)) {
            case STFT_K_FAST: k = pad ? (const void *)k_stft_pipe<PL, XP, PQ, MINB, STFT_K_FAST, true> : (const void *)k_stft_pipe<PL, XP, PQ, MINB, STFT_K_FAST, false>; break;
            case STFT_K_EXACT: k = (const void *)k_stft_pipe<PL, XP, PQ, MINB, STFT_K_EXACT, true>; break;
            default: k = (const void *)k_stft_pipe<PL, XP, PQ, MINB, STFT_K_COMPLEX, true>; break;
        }
        return launch_persistent_raw(k, smem_pipe, PL::T * XP, (sp.frames + XP - 1) / XP, (void *)&sp, s);
    }
    static StftVariant make(const char *name) { return StftVariant{name, PL::N, plan_radices<PL>(), &launch, &launch_pipe, nullptr, {}}; }
};

// span-staged persistent STFT (k_stft_span): SPL = core plan (may differ from the direct kernel's: the one-exchange plans
// the plain r2c kernels use), XS frames per tile.  Returns cudaErrorInvalidConfiguration when the tile's span does not
// fit shared memory (very large hops): the caller falls back to the direct kernel.
template <class PL, int X, int MINB, class SPL, int XS, int SMINB, int SPQ, bool DEFAULT> struct StftSpanLaunchers : StftLaunchers<PL, X, MINB> {
    static cudaError_t launch_span(const StftParams &sp0, cudaStream_t s) {
        StftParams sp = sp0;
        const size_t span = ((size_t)((long)(XS - 1) * sp.hop + sp.wsize) * sizeof(float) + 127) / 128 * 128;
        const size_t smem = stft_work_bytes<SPL, SPQ, XS>() + span + 128 + sizeof(float2) * SPL::N;      // + mbarriers/slots + window
        if (smem > 200u * 1024u) return cudaErrorInvalidConfiguration;
        sp.span_bytes = (int)span;
        const bool pad = sp.wsize < 2 * SPL::N;
        const void *k;
        switch (StftLaunchers<PL, X, MINB>::flavour(sp)) {
            case STFT_K_FAST: k = pad ? (const void *)k_stft_span<SPL, XS, SPQ, SMINB, STFT_K_FAST, true> : (const void *)k_stft_span<SPL, XS, SPQ, SMINB, STFT_K_FAST, false>; break;
            case STFT_K_EXACT: k = (const void *)k_stft_span<SPL, XS, SPQ, SMINB, STFT_K_EXACT, true>; break;
            default: k = (const void *)k_stft_span<SPL, XS, SPQ, SMINB, STFT_K_COMPLEX, true>; break;
        }
        return launch_persistent_dyn(k, smem, SPL::T * XS, (sp.frames + XS - 1) / XS, sp, s);
    }
    static StftVariant make(const char *name) {
        static_assert(SPL::N == PL::N, "same core size");
        return StftVariant{name, PL::N, plan_radices<PL>(), &StftLaunchers<PL, X, MINB>::launch, &StftLaunchers<PL, X, MINB>::launch_pipe, &launch_span, plan_radices<SPL>(), DEFAULT};
    }
};

#define XROWS(T) ((T) >= 256 ? 1 : 256 / (T))

// f32 core plans: the reference's split-core stage structure (radix-4, leading radix-2 for odd log2)
using F32_4 = Plan<4, 1, 0x4>;
using F32_8 = Plan<8, 1, 0x222>;
using F32_16 = Plan<16, 1, 0x44>;
using F32_32 = Plan<32, 2, 0x2, 0x44>;
using F32_64 = Plan<64, 4, 0x4, 0x44>;
using F32_128 = Plan<128, 8, 0x24, 0x44>;
using F32_256 = Plan<256, 16, 0x44, 0x44>;
using F32_512 = Plan<512, 32, 0x2, 0x44, 0x44>;
using F32_1024 = Plan<1024, 64, 0x4, 0x44, 0x44>;
using F32_2048 = Plan<2048, 128, 0x24, 0x44, 0x44>;
using F32_4096 = Plan<4096, 256, 0x44, 0x44, 0x44>;
using F32_8192 = Plan<8192, 512, 0x2, 0x44, 0x44, 0x44>;
// experimental: 32 values per thread (two register blocks per pass), half the threads per transform
using P32_4096_T128 = Plan<4096, 128, 0x44, 0x44, 0x44>;
using P32_2048_T64 = Plan<2048, 64, 0x24, 0x44, 0x44>;
// 64 values per thread: two passes, ONE shared-memory exchange (pad slot per 64 values: tools/bank_sim.py)
using P64_4096 = Plan<4096, 64, 0x444, 0x444>;
using P64_2048 = Plan<2048, 32, 0x244, 0x444>;
using P64_1024 = Plan<1024, 16, 0x44, 0x444>;
// 32 values per thread: two passes, one exchange; a 1024-point transform is ONE warp (no block barriers).
// f32 only (tolerance-based parity lets the stage radices be regrouped as 2,4,4 | 2,4,4).
using P32_1024 = Plan<1024, 32, 0x244, 0x244>;    // pad slot per 32 values
using P32_512 = Plan<512, 16, 0x244, 0x44>;
using P32_8192 = Plan<8192, 256, 0x244, 0x44, 0x44>;   // the same seven stages as F32_8192 in three passes instead of four
// thread-per-row plans for the tile kernels (whole transform in registers)
using T32_32 = Plan<32, 1, 0x244>;
using T32_64 = Plan<64, 1, 0x444>;
// f64 core plans: radix-4 for N = 4^p, radix-2 otherwise (fft_combined.wat:727-732)
using F64_4 = Plan<4, 1, 0x4>;
using F64_8 = Plan<8, 1, 0x222>;
using F64_16 = Plan<16, 1, 0x44>;
using F64_32 = Plan<32, 2, 0x2, 0x2222>;
using T64_32 = Plan<32, 1, 0x22222>;              // thread-per-row (same radix-2 stage sequence)
using F64_64 = Plan<64, 4, 0x4, 0x44>;
using F64_128 = Plan<128, 8, 0x222, 0x2222>;
using F64_256 = Plan<256, 16, 0x44, 0x44>;
using F64_512 = Plan<512, 32, 0x2, 0x2222, 0x2222>;
using F64_1024 = Plan<1024, 64, 0x4, 0x44, 0x44>;
using D32_512 = Plan<512, 16, 0x22222, 0x2222>;      // 32 values per thread: the same nine radix-2 stages, one exchange
using F64_2048 = Plan<2048, 128, 0x222, 0x2222, 0x2222>;
using F64_4096 = Plan<4096, 256, 0x44, 0x44, 0x44>;
using F64_8192 = Plan<8192, 512, 0x2, 0x2222, 0x2222, 0x2222>;

}  // namespace wfb
