// persistent TMA-pipelined r2c / c2r f32 kernels (core size M = N/2)
#include "wfb_registry.h"
namespace wfb {
#define VR(PL, X, MINB, ...) RealPipeLaunchers<float, PL, X, MINB>::make(#PL "_rpipe" #X, __VA_ARGS__)
#define VRQ(PL, X, MINB, PQ, ...) RealPipeLaunchers<float, PL, X, MINB, false, PQ>::make(#PL "_rpipe" #X, __VA_ARGS__)
#define VRTS(PL, X, MINB, PQ, ...) RealPipeLaunchers<float, PL, X, MINB, false, PQ, true>::make(#PL "_rpipe" #X "_ts", __VA_ARGS__)
#define VRC(PL, X, MINB, ...) RealPipeLaunchers<float, PL, X, MINB, true>::make(#PL "_rpipe" #X "_rc", __VA_ARGS__)
const std::vector<Variant> &variants_f32_real_pipe() {
    static const std::vector<Variant> v = {
        VR(F32_64, 32, 2, 30), VR(F32_128, 16, 2, 9, 30), VR(F32_256, 8, 2, 30), VR(F32_512, 4, 2, 30),
        VRC(F32_64, 32, 2, 31, 8), VRC(F32_128, 16, 2, 31, 8), VRC(F32_256, 16, 2, 29, 8),
        // one-exchange plans (32 / 64 values per thread): the real kernels are shared-memory-pipe bound at M >= 512,
        // so dropping an exchange buys 5-20 % there (it buys nothing for c2c at the same M)
        VRQ(P32_512, 4, 2, 16, 31), VRQ(P32_1024, 2, 2, 32, 31), VRQ(P32_1024, 4, 2, 32, 27),
        VRQ(P64_2048, 2, 1, 64, 31), VRQ(P64_4096, 2, 1, 64, 28),
        // results assembled in the stage buffer, one bulk store per tile
        // (profiles/: N = 1024/2048 r2c 88-90 -> 103 % of the HBM peak at burst clocks, 84-86 -> 91-94 % power-capped)
        VRTS(F32_128, 16, 2, 16, 60), VRTS(F32_256, 8, 2, 16, 60), VRTS(P32_512, 4, 2, 16, 60), VRTS(P32_1024, 2, 2, 32, 60),
        VRTS(P64_2048, 2, 1, 64, 20, 60), VRTS(F32_2048, 2, 2, 16, 60, 20), VRTS(F32_4096, 2, 1, 16, 60),
        // single-row (16 KB) tiles at N = 4096, 8192
        // (a row IS the 16 KB tile there; rows are 8 bytes off a 16-byte multiple, see SHIFT / ROW1 in k_real_pipe.
        //  N = 4096: r2c 95 -> 97 %, c2r 87 -> 97 % of the HBM peak at burst clocks; N = 8192: 79 -> 88-89 %)
        RealPipeLaunchers<float, F32_2048, 1, 2, false, 16, true, 1>::make("F32_2048_rpipe1_ts", 61), RealPipeLaunchers<float, F32_4096, 1, 1, false, 16, true, 1>::make("F32_4096_rpipe1_ts", 61),
        RealPipeLaunchers<float, P32_8192, 1, 1, false, 16, true, 1>::make("P32_8192_rpipe1_ts", 61),
        VR(F32_1024, 2, 2, 30), VR(F32_2048, 2, 2, 30, 9), VR(F32_4096, 2, 1, 30),
    };
    return v;
}
}  // namespace wfb
