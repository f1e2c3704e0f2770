// fused STFT kernels: FFT sizes 64..8192 (core M = 32..4096)
#include "wfb_registry.h"
namespace wfb {
// X chosen so a CTA has 128 threads for M <= 1024 (several CTAs per SM interleave their phases)
#define XS(T) ((T) >= 128 ? 1 : 128 / (T))
#define V(PL) StftLaunchers<PL, XS(PL::T), 2>::make(#PL "_stft")
// + the span-staged persistent kernel (core plan, frames per tile, min CTAs/SM, pad quantum)
#define VS(PL, SPL, XSP, SMINB, SPQ) StftSpanLaunchers<PL, XS(PL::T), 2, SPL, XSP, SMINB, SPQ>::make(#PL "_stft+" #SPL "_span" #XSP)
const std::vector<StftVariant> &variants_stft() {
    static const std::vector<StftVariant> v = {
        V(F32_32), V(F32_64),
        // span kernels from N = 256: the plain r2c kernels' plans (one exchange at M = 512, 1024), 128-256 threads per CTA
        VS(F32_128, F32_128, 16, 3, 16), VS(F32_256, F32_256, 8, 3, 16), VS(F32_512, P32_512, 8, 3, 16), VS(F32_1024, P32_1024, 4, 3, 32),
        VS(F32_2048, F32_2048, 2, 2, 16), VS(F32_4096, F32_4096, 1, 2, 16),
    };
    return v;
}
}  // namespace wfb
