// wfb_kernels.cuh -- hand-written sm_100a batched FFT kernels (single HBM pass per transform).
//
// One engine serves the four transforms of the hot path.  Arithmetic follows the reference's
// Stockham autosort DIT stages -- radix-4 with an optional leading radix-2 for the f32 split core
// (modules/fft_split_native_f32.wat:748-888, :710-743), radix-4 for N = 4^p else radix-2 for f64
// (modules/fft_combined.wat:361-474, :486-716) -- so every product `table entry x operand` and
// every add/sub pairing is the one the reference forms (required for the f64 parity bound,
// SURVEY.md F4).  What is B200-specific is the execution: instead of log4(N) passes over memory,
// consecutive reference stages are fused into register-resident "passes" of combined radix up to
// 16 (each thread holds E = N/T complex values), with one padded shared-memory exchange between
// passes, coalesced global loads in the first pass and coalesced global stores in the last, so
// each transform crosses HBM exactly once in each direction.
//
// Index algebra (s = stride, l = groups, N = r*l*s for a stage of radix r):
//   stage reads   r*j*s + t + m*s   (group j < l, lane t < s, input m < r)
//   stage writes  j*s + t + q*N/r   (output q < r)
//   twiddle for input m: W_N^(m*j*N/(r*l))  = stage_table[(m-1)*l + j]
// A pass fusing stages of radices r1..rg (Rp = r1*...*rg) entered with l groups and leaving
// stride s' = N/(l*Rp) gives thread-block b = j*s' + t' the Rp inputs  Rp*s'*j + t' + k*s'
// and the Rp outputs  b + k'*(N/Rp),  k' = mixed-radix digit reversal of the register slot k.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <type_traits>
#include <utility>

namespace wfb {

// Compile-time loop: the body receives std::integral_constant<int, I>, so every index derived from it
// is a constant expression.  (A plain `#pragma unroll` loop left the plan's helper functions to be
// evaluated at run time: ncu showed 2.4x the expected instruction count.)
template <int... Is, class F>
__device__ __forceinline__ void static_for_impl(std::integer_sequence<int, Is...>, F &&f) {
    (f(std::integral_constant<int, Is>{}), ...);
}
template <int N, class F> __device__ __forceinline__ void static_for(F &&f) {
    static_for_impl(std::make_integer_sequence<int, N>{}, static_cast<F &&>(f));
}
#define CIDX(name, tag) constexpr int name = decltype(tag)::value

// ----------------------------------------------------------------------------------------
// compile-time plan description
// ----------------------------------------------------------------------------------------
// A pass is encoded as hex digits, most significant = first stage: 0x44 = radix-4 then radix-4.
__host__ __device__ constexpr int pass_nsub(int code) { int n = 0; while (code) { n++; code >>= 4; } return n; }
__host__ __device__ constexpr int pass_radix(int code, int q) { return (code >> (4 * (pass_nsub(code) - 1 - q))) & 0xF; }
__host__ __device__ constexpr int pass_rp(int code) { int r = 1; while (code) { r *= (code & 0xF); code >>= 4; } return r; }

template <int N_, int T_, int P0, int P1 = 0, int P2 = 0, int P3 = 0, int P4 = 0, int P5 = 0>
struct Plan {
    static constexpr int N = N_;          // complex points per transform
    static constexpr int T = T_;          // threads per transform
    static constexpr int E = N_ / T_;     // complex values held per thread
    static constexpr int MAXP = 6;
    __host__ __device__ static constexpr int code(int p) {
        return p == 0 ? P0 : p == 1 ? P1 : p == 2 ? P2 : p == 3 ? P3 : p == 4 ? P4 : p == 5 ? P5 : 0;
    }
    __host__ __device__ static constexpr int npass() { int n = 0; for (int p = 0; p < MAXP; p++) if (code(p)) n++; return n; }
    // groups on entry to pass p
    __host__ __device__ static constexpr int l_in(int p) { int l = 1; for (int i = 0; i < p; i++) l *= pass_rp(code(i)); return l; }
    // offset (in complex entries) of the table of sub-stage q of pass p
    __host__ __device__ static constexpr int tw_off(int p, int q) {
        int off = 0, l = 1;
        for (int i = 0; i < MAXP; i++) {
            int c = code(i);
            for (int k = 0; k < pass_nsub(c); k++) {
                if (i == p && k == q) return off;
                int r = pass_radix(c, k);
                off += (r - 1) * l;
                l *= r;
            }
        }
        return off;
    }
    __host__ __device__ static constexpr int tw_total() { return tw_off(MAXP, 0); }
    __host__ __device__ static constexpr bool valid() {
        int prod = 1;
        for (int p = 0; p < MAXP; p++) if (code(p)) { prod *= pass_rp(code(p)); if (E % pass_rp(code(p))) return false; }
        return prod == N && N % T == 0;
    }
};

// ----------------------------------------------------------------------------------------
// small vector helpers
// ----------------------------------------------------------------------------------------
template <typename R> struct Vec2;
template <> struct Vec2<float> { using type = float2; };
template <> struct Vec2<double> { using type = double2; };
template <typename R> using vec2 = typename Vec2<R>::type;

template <typename R> __device__ __forceinline__ vec2<R> mk2(R a, R b) { vec2<R> v; v.x = a; v.y = b; return v; }
template <typename R> __device__ __forceinline__ vec2<R> cmul(vec2<R> w, vec2<R> v) {
    // (wr*vr - wi*vi, wr*vi + wi*vr): fft_split_native_f32.wat:826-837 / fft_combined.wat:417-419
    return mk2<R>(w.x * v.x - w.y * v.y, w.x * v.y + w.y * v.x);
}
template <typename R> __device__ __forceinline__ vec2<R> cadd(vec2<R> a, vec2<R> b) { return mk2<R>(a.x + b.x, a.y + b.y); }
template <typename R> __device__ __forceinline__ vec2<R> csub(vec2<R> a, vec2<R> b) { return mk2<R>(a.x - b.x, a.y - b.y); }

// shared-memory slot of logical element p (units of one complex value).  One pad slot per PADQ
// elements keeps both the contiguous writes (tid + e*T) and the strided gathers of the later
// passes (Rp*s'*j + t' + k*s') conflict-free; verified by tools/bank_sim.py for every plan.
template <int PADQ> __device__ __forceinline__ constexpr int pad_idx(int p) { return PADQ ? p + (p / PADQ) : p; }
template <int PADQ> __host__ __device__ constexpr int padded_size(int n) { return PADQ ? n + (n / PADQ) : n; }

// mixed-radix helpers over the sub-stages of one pass
// hi enumerates the already-transformed (more significant) digits before sub-stage q.
__host__ __device__ constexpr int hi_count(int code, int q) { int h = 1; for (int a = 0; a < q; a++) h *= pass_radix(code, a); return h; }
// c(hi): output-group offset contributed by those digits = sum m'_a * prod_{h<a} r_h
__host__ __device__ constexpr int hi_to_c(int code, int q, int hi) {
    // digits of hi: m'_0 most significant ... m'_{q-1} least significant
    int c = 0, w = hi_count(code, q);
    int mult = 1;
    for (int a = 0; a < q; a++) {
        int r = pass_radix(code, a);
        w /= r;
        int d = (hi / w) % r;
        c += d * mult;
        mult *= r;
    }
    return c;
}
// final register slot k -> output digit-reversed index k'
__host__ __device__ constexpr int slot_to_out(int code, int k) {
    int g = pass_nsub(code), w = pass_rp(code), mult = 1, out = 0;
    for (int a = 0; a < g; a++) {
        int r = pass_radix(code, a);
        w /= r;
        int d = (k / w) % r;
        out += d * mult;
        mult *= r;
    }
    return out;
}

// ----------------------------------------------------------------------------------------
// one fused pass over the E register-resident values of a thread
// ----------------------------------------------------------------------------------------
template <typename R, class PL, int P, bool INV>
__device__ __forceinline__ void run_pass(vec2<R> (&x)[PL::E], const vec2<R> *__restrict__ tw, int tid) {
    constexpr int CODE = PL::code(P);
    constexpr int RP = pass_rp(CODE);
    constexpr int NB = PL::E / RP;                 // register blocks per thread
    constexpr int L_IN = PL::l_in(P);
    constexpr int SP = PL::N / (L_IN * RP);        // stride on exit (s')
    constexpr int G = pass_nsub(CODE);
    static_for<NB>([&](auto I_) {
        CIDX(i, I_);
        const int b = tid + i * PL::T;
        const int j = (L_IN == 1) ? 0 : b / SP;
        static_for<G>([&](auto Q_) {
            CIDX(q, Q_);
            constexpr int r = pass_radix(CODE, q);
            constexpr int H = hi_count(CODE, q);
            constexpr int w = RP / (H * r);            // slot stride of this digit
            constexpr int lq = L_IN * H;               // groups of this stage
            constexpr int off = PL::tw_off(P, q);
            static_for<H>([&](auto HI_) {
                CIDX(hi, HI_);
                constexpr int c = hi_to_c(CODE, q, hi);
                constexpr bool unit = (L_IN == 1 && c == 0);     // group 0: W^0 = 1
                vec2<R> w1 = mk2<R>(R(1), R(0)), w2 = w1, w3 = w1;
                if constexpr (!unit) {
                    const int jq = j + L_IN * c;
                    w1 = __ldg(tw + off + jq);
                    if constexpr (r == 4) {
                        w2 = __ldg(tw + off + lq + jq);
                        w3 = __ldg(tw + off + 2 * lq + jq);
                    }
                }
                static_for<w>([&](auto LO_) {
                    CIDX(lo, LO_);
                    constexpr int k0 = hi * (w * r) + lo;
                    if constexpr (r == 2) {
                        vec2<R> &A = x[i + NB * k0], &B = x[i + NB * (k0 + w)];
                        const vec2<R> wb = unit ? B : cmul<R>(w1, B);
                        const vec2<R> a = A;
                        A = cadd<R>(a, wb);
                        B = csub<R>(a, wb);
                    } else {
                        vec2<R> &A = x[i + NB * k0], &B = x[i + NB * (k0 + w)];
                        vec2<R> &C = x[i + NB * (k0 + 2 * w)], &D = x[i + NB * (k0 + 3 * w)];
                        const vec2<R> wb = unit ? B : cmul<R>(w1, B);
                        const vec2<R> wc = unit ? C : cmul<R>(w2, C);
                        const vec2<R> wd = unit ? D : cmul<R>(w3, D);
                        const vec2<R> t0 = cadd<R>(A, wc), t1 = csub<R>(A, wc);
                        const vec2<R> t2 = cadd<R>(wb, wd), t3 = csub<R>(wb, wd);
                        A = cadd<R>(t0, t2);
                        C = csub<R>(t0, t2);
                        // forward: out1 = t1 - i*t3, out3 = t1 + i*t3; the inverse swaps them
                        // (fft_split_native_f32.wat:785-788)
                        const vec2<R> m1 = mk2<R>(t1.x + t3.y, t1.y - t3.x);
                        const vec2<R> m3 = mk2<R>(t1.x - t3.y, t1.y + t3.x);
                        B = INV ? m3 : m1;
                        D = INV ? m1 : m3;
                    }
                });
            });
        });
    });
}

// after a pass, register slot (i, k) holds logical element  tid + (i + NB*k')*T,  k' = slot_to_out(k)
template <class PL, int P> __host__ __device__ constexpr int out_elem(int i, int k) {
    return i + (PL::E / pass_rp(PL::code(P))) * slot_to_out(PL::code(P), k);
}

// ----------------------------------------------------------------------------------------
// barriers scoped to the threads of one transform
// ----------------------------------------------------------------------------------------
template <int T, int X> __device__ __forceinline__ void sync_transform(int xi) {
    if constexpr (T <= 32) {
        __syncwarp();
    } else if constexpr (X == 1) {
        __syncthreads();
    } else {
        asm volatile("bar.sync %0, %1;" ::"r"(xi + 1), "n"(T) : "memory");
    }
}

// exchange: registers -> smem (layout of pass P's outputs) -> registers (layout of pass P+1's inputs)
template <typename R, class PL, int P, int PADQ, int X>
__device__ __forceinline__ void exchange(vec2<R> (&x)[PL::E], vec2<R> *sm, int tid, int xi, bool need_pre_sync) {
    constexpr int CODE = PL::code(P);
    constexpr int RP = pass_rp(CODE);
    constexpr int NB = PL::E / RP;
    if (need_pre_sync) sync_transform<PL::T, X>(xi);   // everyone finished reading the previous contents
    static_for<PL::E>([&](auto S_) {
        CIDX(slot, S_);
        constexpr int i = slot % NB, k = slot / NB;
        constexpr int e = out_elem<PL, P>(i, k);
        sm[pad_idx<PADQ>(tid + e * PL::T)] = x[slot];
    });
    sync_transform<PL::T, X>(xi);
    constexpr int CODE2 = PL::code(P + 1);
    constexpr int RP2 = pass_rp(CODE2);
    constexpr int NB2 = PL::E / RP2;
    constexpr int SP2 = PL::N / (PL::l_in(P + 1) * RP2);
    static_for<NB2>([&](auto I_) {
        CIDX(i, I_);
        const int b = tid + i * PL::T;
        const int j = b / SP2, t = b % SP2;
        const int base = RP2 * SP2 * j + t;
        static_for<RP2>([&](auto K_) {
            CIDX(k, K_);
            x[i + NB2 * k] = sm[pad_idx<PADQ>(base + k * SP2)];
        });
    });
}

// all passes after the first-pass inputs are in registers; leaves the last pass's outputs in x
template <typename R, class PL, int PADQ, int X, bool INV, int P = 0>
__device__ __forceinline__ void run_all(vec2<R> (&x)[PL::E], const vec2<R> *__restrict__ tw, vec2<R> *sm, int tid,
                                        int xi, bool smem_dirty) {
    run_pass<R, PL, P, INV>(x, tw, tid);
    if constexpr (P + 1 < PL::npass()) {
        exchange<R, PL, P, PADQ, X>(x, sm, tid, xi, smem_dirty || P > 0);
        run_all<R, PL, PADQ, X, INV, P + 1>(x, tw, sm, tid, xi, true);
    }
}

// ----------------------------------------------------------------------------------------
// kernel parameters
// ----------------------------------------------------------------------------------------
struct KParams {
    const void *in0, *in1;     // input planes
    void *out0, *out1;         // output planes
    const void *tw;            // stage tables for this direction (complex entries)
    const void *rtw;           // W_Nreal^k, k = 0..M, for the real transforms (complex entries)
    long batch;
    double scale;              // applied on store (1/N for the inverse c2c)
};

enum IoMode { IO_SPLIT = 0, IO_INTERLEAVED = 1 };

// streaming global accesses: every payload byte is touched exactly once
template <typename V> __device__ __forceinline__ V ld_stream(const V *p) { return __ldcs(p); }
template <typename V> __device__ __forceinline__ void st_stream(V *p, V v) { __stcs(p, v); }

// ----------------------------------------------------------------------------------------
// Transform 1 and 3 (f32) / 4 (f64): batched c2c, split or interleaved I/O
// ----------------------------------------------------------------------------------------
template <typename R, class PL, int X, int PADQ, int IO, bool INV, int MINB>
__global__ void __launch_bounds__(PL::T *X, MINB) k_c2c(KParams p) {
    static_assert(PL::valid(), "plan does not factor N");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int LAST = PL::npass() - 1;
    const int xi = threadIdx.x / PL::T, tid = threadIdx.x % PL::T;
    const long row = (long)blockIdx.x * X + xi;
    const bool active = row < p.batch;
    vec2<R> *sm = reinterpret_cast<vec2<R> *>(smem_raw) + (size_t)xi * padded_size<PADQ>(PL::N);
    const vec2<R> *tw = reinterpret_cast<const vec2<R> *>(p.tw);
    vec2<R> x[PL::E];

    if (active) {
        if constexpr (IO == IO_SPLIT) {
            const R *re = reinterpret_cast<const R *>(p.in0) + row * PL::N;
            const R *im = reinterpret_cast<const R *>(p.in1) + row * PL::N;
#pragma unroll
            for (int e = 0; e < PL::E; e++) x[e].x = ld_stream(re + tid + e * PL::T);
#pragma unroll
            for (int e = 0; e < PL::E; e++) x[e].y = ld_stream(im + tid + e * PL::T);
        } else {
            const vec2<R> *z = reinterpret_cast<const vec2<R> *>(p.in0) + row * PL::N;
#pragma unroll
            for (int e = 0; e < PL::E; e++) x[e] = ld_stream(z + tid + e * PL::T);
        }
    } else {
#pragma unroll
        for (int e = 0; e < PL::E; e++) x[e] = mk2<R>(R(0), R(0));
    }

    run_all<R, PL, PADQ, X, INV>(x, tw, sm, tid, xi, false);

    if (active) {
        constexpr int RP = pass_rp(PL::code(LAST));
        constexpr int NB = PL::E / RP;
        const R sc = (R)p.scale;
        if constexpr (IO == IO_SPLIT) {
            R *re = reinterpret_cast<R *>(p.out0) + row * PL::N;
            R *im = reinterpret_cast<R *>(p.out1) + row * PL::N;
            static_for<PL::E>([&](auto S_) {
                CIDX(slot, S_);
                constexpr int e = out_elem<PL, LAST>(slot % NB, slot / NB);
                st_stream(re + tid + e * PL::T, INV ? x[slot].x * sc : x[slot].x);
            });
            static_for<PL::E>([&](auto S_) {
                CIDX(slot, S_);
                constexpr int e = out_elem<PL, LAST>(slot % NB, slot / NB);
                st_stream(im + tid + e * PL::T, INV ? x[slot].y * sc : x[slot].y);
            });
        } else {
            vec2<R> *z = reinterpret_cast<vec2<R> *>(p.out0) + row * PL::N;
            static_for<PL::E>([&](auto S_) {
                CIDX(slot, S_);
                constexpr int e = out_elem<PL, LAST>(slot % NB, slot / NB);
                vec2<R> v = x[slot];
                if (INV) v = mk2<R>(v.x * sc, v.y * sc);
                st_stream(z + tid + e * PL::T, v);
            });
        }
    }
}

// ----------------------------------------------------------------------------------------
// Transform 2 (f32) / 4-real (f64): batched r2c.  PL describes the M = N/2 point complex core.
// The even/odd deinterleave is the float2 view of the packed input (fused into the first-stage
// loads, cf. $stage_r4_first_fused :1376-1459); the Hermitian post-process runs out of shared
// memory straight into the coalesced spectrum store (cf. $rfft_postprocess_split :1471-1559,
// fft_real_combined.wat:455-592).
// ----------------------------------------------------------------------------------------
template <typename R> struct RealPost;

// f32 flavour: one twiddle W^k serves both X[k] and X[M-k]
template <> struct RealPost<float> {
    __device__ static __forceinline__ void pair(float2 z, float2 zm, float2 w, float2 wm, float2 &xk, float2 &xm) {
        (void)wm;
        float gr = z.x + zm.x, gi = z.y - zm.y, hr = z.y + zm.y, hi = zm.x - z.x;
        float tr = w.x * hr - w.y * hi, ti = w.x * hi + w.y * hr;
        xk = make_float2(0.5f * (gr + tr), 0.5f * (gi + ti));
        xm = make_float2(0.5f * (gr - tr), 0.5f * (ti - gi));
    }
    // X[M/2]: the reference's last vector iteration stores the mirrored form last (:1527-1545);
    // the M = 32 fused ending stores conj(Z[M/2]) (:2710)
    __device__ static __forceinline__ float2 middle(float2 z, float2 w, int m) {
        if (m == 32) return make_float2(z.x, -z.y);
        float2 a, b;
        pair(z, z, w, w, a, b);
        return b;
    }
};
// f64 flavour: T[k] and T[M-k] are separately tabulated (fft_real_combined.wat:502-503,533-534)
template <> struct RealPost<double> {
    __device__ static __forceinline__ double2 one(double2 z, double2 zm, double2 w) {
        double sr = z.x + zm.x, si = z.y - zm.y, dr = z.x - zm.x, di = z.y + zm.y;
        double wdr = w.y * dr + w.x * di, wdi = w.y * di - w.x * dr;
        return make_double2(0.5 * (sr + wdr), 0.5 * (si + wdi));
    }
    __device__ static __forceinline__ void pair(double2 z, double2 zm, double2 w, double2 wm, double2 &xk, double2 &xm) {
        xk = one(z, zm, w);
        xm = one(zm, z, wm);
    }
    __device__ static __forceinline__ double2 middle(double2 z, double2 w, int) {
        // sum = (2 re, 0), diff = (0, 2 im)  (fft_real_combined.wat:1031-1050)
        double sr = 2.0 * z.x, si = 0.0, dr = 0.0, di = 2.0 * z.y;
        double wdr = w.y * dr + w.x * di, wdi = w.y * di - w.x * dr;
        return make_double2(0.5 * (sr + wdr), 0.5 * (si + wdi));
    }
};

template <typename R, class PL, int X, int PADQ, int MINB>
__global__ void __launch_bounds__(PL::T *X, MINB) k_r2c(KParams p) {
    static_assert(PL::valid(), "plan does not factor N");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int M = PL::N;
    constexpr int LAST = PL::npass() - 1;
    const int xi = threadIdx.x / PL::T, tid = threadIdx.x % PL::T;
    const long row = (long)blockIdx.x * X + xi;
    const bool active = row < p.batch;
    vec2<R> *sm = reinterpret_cast<vec2<R> *>(smem_raw) + (size_t)xi * padded_size<PADQ>(M);
    const vec2<R> *tw = reinterpret_cast<const vec2<R> *>(p.tw);
    const vec2<R> *rtw = reinterpret_cast<const vec2<R> *>(p.rtw);
    vec2<R> x[PL::E];

    if (active) {
        const vec2<R> *z = reinterpret_cast<const vec2<R> *>(p.in0) + row * M;   // z[j] = x[2j] + i x[2j+1]
#pragma unroll
        for (int e = 0; e < PL::E; e++) x[e] = ld_stream(z + tid + e * PL::T);
    } else {
#pragma unroll
        for (int e = 0; e < PL::E; e++) x[e] = mk2<R>(R(0), R(0));
    }

    run_all<R, PL, PADQ, X, false>(x, tw, sm, tid, xi, false);

    // Z -> shared memory in natural order
    {
        constexpr int RP = pass_rp(PL::code(LAST));
        constexpr int NB = PL::E / RP;
        if (PL::npass() > 1) sync_transform<PL::T, X>(xi);
        static_for<PL::E>([&](auto S_) {
            CIDX(slot, S_);
            constexpr int e = out_elem<PL, LAST>(slot % NB, slot / NB);
            sm[pad_idx<PADQ>(tid + e * PL::T)] = x[slot];
        });
        sync_transform<PL::T, X>(xi);
    }
    if (!active) return;

    vec2<R> *out = reinterpret_cast<vec2<R> *>(p.out0) + row * (M + 1);
    // pairs (k, M-k), k = tid + i*T over 0 .. M/2-1; k = 0 is DC/Nyquist; thread 0 adds k = M/2
    constexpr int HALF = M / 2;
    constexpr int PER = (HALF + PL::T - 1) / PL::T;
#pragma unroll
    for (int i = 0; i < PER; i++) {
        const int k = tid + i * PL::T;
        if (k >= HALF) break;
        if (k == 0) {
            vec2<R> z0 = sm[0];
            st_stream(out, mk2<R>(z0.x + z0.y, R(0)));
            st_stream(out + M, mk2<R>(z0.x - z0.y, R(0)));
            vec2<R> zh = sm[pad_idx<PADQ>(HALF)];
            st_stream(out + HALF, RealPost<R>::middle(zh, __ldg(rtw + HALF), M));
        } else {
            vec2<R> z = sm[pad_idx<PADQ>(k)], zm = sm[pad_idx<PADQ>(M - k)];
            vec2<R> xk, xm;
            RealPost<R>::pair(z, zm, __ldg(rtw + k), __ldg(rtw + (M - k)), xk, xm);
            st_stream(out + k, xk);
            st_stream(out + (M - k), xm);
        }
    }
}

// ----------------------------------------------------------------------------------------
// c2r: Hermitian pre-process (scale 0.5/M folded in, $irfft_preprocess_split :1656-1748) fused
// into the loads, inverse M-point core, re-interleave on store ($stage_r4_s1_inv_fused :1756-1932).
// The f64 variant has no reference counterpart (extension, parity unpinned) and uses the same
// formulas in double.
// ----------------------------------------------------------------------------------------
template <typename R, class PL, int X, int PADQ, int MINB>
__global__ void __launch_bounds__(PL::T *X, MINB) k_c2r(KParams p) {
    static_assert(PL::valid(), "plan does not factor N");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int M = PL::N;
    constexpr int LAST = PL::npass() - 1;
    const int xi = threadIdx.x / PL::T, tid = threadIdx.x % PL::T;
    const long row = (long)blockIdx.x * X + xi;
    const bool active = row < p.batch;
    vec2<R> *sm = reinterpret_cast<vec2<R> *>(smem_raw) + (size_t)xi * padded_size<PADQ>(M);
    const vec2<R> *tw = reinterpret_cast<const vec2<R> *>(p.tw);
    const vec2<R> *rtw = reinterpret_cast<const vec2<R> *>(p.rtw);
    vec2<R> x[PL::E];

    constexpr int HALF = M / 2;
    constexpr int PER = (HALF + PL::T - 1) / PL::T;
    const R sc = R(0.5) / R(M);
    if (active) {
        const vec2<R> *in = reinterpret_cast<const vec2<R> *>(p.in0) + row * (M + 1);
#pragma unroll
        for (int i = 0; i < PER; i++) {
            const int k = tid + i * PL::T;
            if (k >= HALF) break;
            if (k == 0) {
                const R x0 = ld_stream(in).x, xm = ld_stream(in + M).x;      // real parts only (:1679-1684)
                sm[0] = mk2<R>((x0 + xm) * sc, (x0 - xm) * sc);
            }
            const int kk = (k == 0) ? HALF : k;     // thread 0's slot k = 0 also covers the self-paired k = M/2
            const vec2<R> a = ld_stream(in + kk), b = ld_stream(in + (M - kk));
            const vec2<R> w = __ldg(rtw + kk);
            const R gr = a.x + b.x, gi = a.y - b.y, ur = a.x - b.x, ui = a.y + b.y;
            const R hr = w.x * ur + w.y * ui, hi = w.x * ui - w.y * ur;
            // forward store first, mirrored second: at k = M/2 the mirrored form survives (:1722-1740)
            sm[pad_idx<PADQ>(kk)] = mk2<R>(sc * (gr - hi), sc * (gi + hr));
            sm[pad_idx<PADQ>(M - kk)] = mk2<R>(sc * (gr + hi), sc * (hr - gi));
        }
    }
    sync_transform<PL::T, X>(xi);
#pragma unroll
    for (int e = 0; e < PL::E; e++) x[e] = sm[pad_idx<PADQ>(tid + e * PL::T)];

    run_all<R, PL, PADQ, X, true>(x, tw, sm, tid, xi, true);

    if (active) {
        constexpr int RP = pass_rp(PL::code(LAST));
        constexpr int NB = PL::E / RP;
        vec2<R> *z = reinterpret_cast<vec2<R> *>(p.out0) + row * M;
        static_for<PL::E>([&](auto S_) {
            CIDX(slot, S_);
            constexpr int e = out_elem<PL, LAST>(slot % NB, slot / NB);
            st_stream(z + tid + e * PL::T, x[slot]);
        });
    }
}

}  // namespace wfb
