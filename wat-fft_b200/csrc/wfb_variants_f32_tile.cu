// thread-per-row tile kernels for N <= 64
#include "wfb_registry.h"
namespace wfb {
#define V(PL, X, MINB, PRIO) TileLaunchers<PL, X, MINB>::make(#PL "_tile", PRIO)
#define VX(PL, X, MINB, PRIO) TileLaunchers<PL, X, MINB>::make(#PL "_tile" #X, PRIO)
#define VRX(PL, X, MINB, ...) RealTileLaunchers<PL, X, MINB>::make(#PL "_rtile" #X, __VA_ARGS__)
#define VR(PL, X, MINB, PRIO) RealTileLaunchers<PL, X, MINB>::make(#PL "_rtile", PRIO)
const std::vector<Variant> &variants_f32_tile() {
    static const std::vector<Variant> v = {
        V(F32_4, 256, 2, 50), V(F32_8, 256, 2, 50), V(F32_16, 256, 2, 50), V(T32_32, 128, 2, 50), V(T32_64, 128, 1, 50),
        VX(T32_64, 32, 1, 52), VX(T32_32, 64, 2, 52), VX(F32_16, 128, 2, 51),
        VR(F32_16, 256, 2, 50), VR(T32_32, 128, 2, 50), VRX(F32_4, 128, 2, 51), VRX(F32_8, 128, 2, 51), VRX(T32_64, 32, 1, 26, 32), VRX(T32_32, 32, 2, 51), VRX(F32_16, 64, 2, 51),
    };
    return v;
}
}  // namespace wfb
