// thread-per-row tile kernels for N <= 64
#include "wfb_registry.h"
namespace wfb {
#define V(PL, X, MINB, PRIO) TileLaunchers<PL, X, MINB>::make(#PL "_tile", PRIO)
#define VX(PL, X, MINB, PRIO) TileLaunchers<PL, X, MINB>::make(#PL "_tile" #X, PRIO)
#define VTP(PL, X, MINB, ...) TilePipeLaunchers<PL, X, MINB>::make(#PL "_tpipe" #X, __VA_ARGS__)
#define VRTP(PL, X, MINB, ...) RealTilePipeLaunchers<PL, X, MINB>::make(#PL "_rtpipe" #X, __VA_ARGS__)
#define VRX(PL, X, MINB, ...) RealTileLaunchers<PL, X, MINB>::make(#PL "_rtile" #X, __VA_ARGS__)
#define VR(PL, X, MINB, PRIO) RealTileLaunchers<PL, X, MINB>::make(#PL "_rtile", PRIO)
const std::vector<Variant> &variants_f32_tile() {
    // *_tpipe / *_rtpipe: persistent, fully TMA-fed (16 KB tiles); defaults for N = 16..64 c2c and N = 64, 128 real
    // (profiles/: 1.00-1.04 of the measured HBM peak, also under the power cap; the staged *_tile kernels: 0.82-1.02)
    static const std::vector<Variant> v = {
        V(F32_4, 256, 2, 50), V(F32_8, 256, 2, 50), V(F32_16, 256, 2, 50), V(T32_32, 128, 2, 50), V(T32_64, 128, 1, 50),
        VX(T32_64, 32, 1, 52), VX(T32_32, 64, 2, 52), VX(F32_16, 128, 2, 51),
        VTP(T32_64, 32, 1, 55), VTP(T32_32, 64, 2, 55), VTP(F32_16, 128, 2, 55), VTP(T32_64, 64, 1, 56, -1, 54),
        VRTP(T32_64, 32, 1, 55), VRTP(T32_32, 64, 2, 55),
        VR(F32_16, 256, 2, 50), VR(T32_32, 128, 2, 50), VRX(F32_4, 128, 2, 51), VRX(F32_8, 128, 2, 51), VRX(T32_64, 32, 1, 26, 32), VRX(T32_32, 32, 2, 51), VRX(F32_16, 64, 2, 51),
    };
    return v;
}
}  // namespace wfb
