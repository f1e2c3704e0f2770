// fused STFT kernels: FFT sizes 64..8192 (core M = 32..4096)
#include "wfb_registry.h"
namespace wfb {
// X chosen so a CTA has 128 threads for M <= 1024 (several CTAs per SM interleave their phases)
#define XS(T) ((T) >= 128 ? 1 : 128 / (T))
#define V(PL) StftLaunchers<PL, XS(PL::T), 2>::make(#PL "_stft")
// + the span-staged persistent kernel (core plan, frames per tile, min CTAs/SM, pad quantum)
#define VS(PL, SPL, XSP, SMINB, SPQ, DFLT) StftSpanLaunchers<PL, XS(PL::T), 2, SPL, XSP, SMINB, SPQ, DFLT>::make(#PL "_stft+" #SPL "_span" #XSP)
const std::vector<StftVariant> &variants_stft() {
    static const std::vector<StftVariant> v = {
        V(F32_32), V(F32_64),
        // span kernels from N = 256: the plain r2c kernels' plans (one exchange at M = 512, 1024), 128-256 threads per CTA.
        // Default where measured faster than the direct kernel at hop = N/4 (frames/s, burst clocks, span vs direct vs the plain
        // r2c's rows/s):  N = 1024: 818 vs 769 (r2c 828);  N = 2048: 406 vs 341 (r2c 415);  N = 256: 2934 vs 3093;  N = 512: 1636 vs
        // 1718;  N = 4096: 154 vs 193;  N = 8192: 71 vs 72 -- the one-exchange cores are what the span kernel buys
        VS(F32_128, F32_128, 16, 3, 16, false), VS(F32_256, F32_256, 8, 3, 16, false), VS(F32_512, P32_512, 8, 4, 16, true),
        VS(F32_1024, P32_1024, 4, 4, 32, true), VS(F32_2048, F32_2048, 2, 2, 16, false), VS(F32_4096, F32_4096, 2, 1, 16, false),
    };
    return v;
}
}  // namespace wfb
