// fused STFT kernels: FFT sizes 64..8192 (core M = 32..4096)
#include "wfb_registry.h"
namespace wfb {
// X chosen so a CTA has 128 threads for M <= 1024 (several CTAs per SM interleave their phases)
#define XS(T) ((T) >= 128 ? 1 : 128 / (T))
#define V(PL) StftLaunchers<PL, XS(PL::T), 2>::make(#PL "_stft")
const std::vector<StftVariant> &variants_stft() {
    static const std::vector<StftVariant> v = {
        // (the one-exchange P32 cores of the plain r2c kernels were measured here too: no gain -- this kernel is bound by
        //  the frame gather out of L2 and the per-bin dB arithmetic, not by the exchanges)
        V(F32_32), V(F32_64), V(F32_128), V(F32_256), V(F32_512), V(F32_1024), V(F32_2048), V(F32_4096),
    };
    return v;
}
}  // namespace wfb
