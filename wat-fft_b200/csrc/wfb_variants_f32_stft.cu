// fused STFT kernels: FFT sizes 64..8192 (core M = 32..4096)
#include "wfb_registry.h"
namespace wfb {
#define V(PL) StftLaunchers<PL, XROWS(PL::T), 2>::make(#PL "_stft")
const std::vector<StftVariant> &variants_stft() {
    static const std::vector<StftVariant> v = {
        V(F32_32), V(F32_64), V(F32_128), V(F32_256), V(F32_512), V(F32_1024), V(F32_2048), V(F32_4096),
    };
    return v;
}
}  // namespace wfb
