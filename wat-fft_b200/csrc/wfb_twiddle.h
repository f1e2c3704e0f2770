// wfb_twiddle.h -- host-side, reference-exact twiddle generation (product code).
//
// The reference builds its W_n^k tables with Taylor-series sin/cos
// (modules/fft_split_native_f32.wat:68-143 for f32, modules/fft_combined.wat:43-106 for f64).
// Those tables are only ~6e-7 (f32) / ~6e-11 (f64) accurate, so to stay within
// 1e-14*log2(N) of the reference's f64 output the GPU kernels must multiply by the SAME
// table entries.  This header restates the trig bit-for-bit (plain IEEE ops, no FMA: the
// translation unit that includes it is compiled with -ffp-contract=off / --fmad=false for
// host code) and derives the per-stage tables the kernels read.
#pragma once
#include <cmath>
#include <cstdint>
#include <vector>

namespace wfb {

template <typename R> struct TrigConst;
template <> struct TrigConst<float> {
    static constexpr float pi = 3.14159265358979323846f, half_pi = 1.5707963267948966f;
    static constexpr int terms = 5;
};
template <> struct TrigConst<double> {
    static constexpr double pi = 3.14159265358979323846, half_pi = 1.5707963267948966;
    static constexpr int terms = 7;
};

// volatile stores defeat host-compiler FMA contraction regardless of flags.
template <typename R> static inline R rnd(R v) { volatile R t = v; return t; }

template <typename R> R ref_sin(R x) {
    using C = TrigConst<R>;
    if (x < -C::pi) x = rnd<R>(x + rnd<R>(R(2) * C::pi));
    if (x > C::pi) x = rnd<R>(x - rnd<R>(R(2) * C::pi));
    if (x > C::half_pi) x = rnd<R>(C::pi - x);
    if (x < -C::half_pi) x = rnd<R>(-C::pi - x);
    const R d[7] = {R(-6), R(-20), R(-42), R(-72), R(-110), R(-156), R(-210)};
    R x2 = rnd<R>(x * x), sum = x, term = x;
    for (int i = 0; i < C::terms; i++) {
        term = rnd<R>(term * rnd<R>(x2 / d[i]));
        sum = rnd<R>(sum + term);
    }
    return sum;
}

template <typename R> R ref_cos(R x) {
    using C = TrigConst<R>;
    R sign = R(1);
    if (x < -C::pi) x = rnd<R>(x + rnd<R>(R(2) * C::pi));
    if (x > C::pi) x = rnd<R>(x - rnd<R>(R(2) * C::pi));
    if (x > C::half_pi) { x = rnd<R>(C::pi - x); sign = R(-1); }
    if (x < -C::half_pi) { x = rnd<R>(C::pi + x); sign = R(-1); }
    const R d[7] = {R(-2), R(-12), R(-30), R(-56), R(-90), R(-132), R(-182)};
    R x2 = rnd<R>(x * x), sum = R(1), term = R(1);
    for (int i = 0; i < C::terms; i++) {
        term = rnd<R>(term * rnd<R>(x2 / d[i]));
        sum = rnd<R>(sum + term);
    }
    return rnd<R>(sum * sign);
}

enum TwFlavour {
    TW_F32_SPLIT = 0,   // angle = f32(-6.283185307)/f32(n) * f32(k)   (fft_split_native_f32.wat:159,167)
    TW_F32_DUAL = 1,    // angle = f32(k) * ((-2*PI)/n)                (fft_stockham_f32_dual.wat:129-141)
    TW_F64 = 2,         // angle = f64(k) * ((-2*PI)/n)                (fft_combined.wat:114-124)
    TW_EXACT = 3        // correctly rounded cos/sin: the hard-coded codelet constants
};

// W_n^k for k < count, reference-exact.
template <typename R>
void base_twiddles(int flavour, int n, int count, std::vector<R> &re, std::vector<R> &im) {
    re.resize(count); im.resize(count);
    if (flavour == TW_EXACT) {
        for (int k = 0; k < count; k++) {
            double a = -2.0 * M_PI * double(k) / double(n);
            double c = std::cos(a), s = std::sin(a);
            if ((4LL * k) % n == 0) {
                int m = int((4LL * k) / n % 4);
                c = (m == 0) ? 1.0 : (m == 2) ? -1.0 : 0.0;
                s = (m == 1) ? -1.0 : (m == 3) ? 1.0 : 0.0;
            } else if ((8LL * k) % n == 0) {
                const double h = 0.7071067811865476;
                c = (c > 0) ? h : -h; s = (s > 0) ? h : -h;
            }
            re[k] = R(c); im[k] = R(s);
        }
        return;
    }
    R step;
    if (flavour == TW_F32_SPLIT) step = rnd<R>(R(-6.283185307f) / R(uint32_t(n)));
    else step = rnd<R>(rnd<R>(R(-2) * TrigConst<R>::pi) / R(uint32_t(n)));
    for (int k = 0; k < count; k++) {
        R a = (flavour == TW_F32_SPLIT) ? rnd<R>(step * R(uint32_t(k))) : rnd<R>(R(uint32_t(k)) * step);
        re[k] = ref_cos<R>(a);
        im[k] = ref_sin<R>(a);
    }
}

// Per-stage tables, concatenated in stage order.  Stage q (radix r, l groups) contributes
// (r-1)*l interleaved (re,im) entries at index (m-1)*l + j holding W_n^(m*j*n/(r*l)) -- the very
// entries the reference looks up (fft_split_native_f32.wat:225-252 $build_r4_tables;
// fft_combined.wat:395-409).  `inverse` conjugates (the reference's STAGE_TW_INV, :188-189).
template <typename R>
void stage_tables(const std::vector<int> &radices, int n, const std::vector<R> &bre, const std::vector<R> &bim,
                  bool inverse, std::vector<R> &out) {
    out.clear();
    int l = 1;
    for (int r : radices) {
        long step = n / (long(r) * l);
        for (int m = 1; m < r; m++)
            for (int j = 0; j < l; j++) {
                long idx = (long(m) * j * step) % n;
                out.push_back(bre[idx]);
                out.push_back(inverse ? -bim[idx] : bim[idx]);
            }
        l *= r;
    }
}

}  // namespace wfb
