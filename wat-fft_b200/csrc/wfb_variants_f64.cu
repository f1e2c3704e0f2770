// f64 kernels: c2c interleaved, r2c, c2r (extension)
#include "wfb_registry.h"
namespace wfb {
#define V(PL, MINB, PRIO) Launchers<double, PL, XROWS(PL::T), MINB, false>::make(#PL, PRIO)
#define VTP(PL, X, MINB, PRIO) TilePipeLaunchers<PL, X, MINB, double>::make(#PL "_tpipe" #X, PRIO)
const std::vector<Variant> &variants_f64() {
    static const std::vector<Variant> v = {
        // persistent, fully TMA-fed thread-per-row c2c (see k_c2c_tpipe): 16 KB tiles
        VTP(F64_4, 256, 2, 40), VTP(F64_8, 128, 2, 40), VTP(F64_16, 64, 2, 40), VTP(T64_32, 32, 1, 40),
        // ... and r2c / c2r at N = 8..64 (k_real_tpipe)
        RealTilePipeLaunchers<F64_4, 256, 2, double>::make("F64_4_rtpipe256", 40), RealTilePipeLaunchers<F64_8, 128, 2, double>::make("F64_8_rtpipe128", 40),
        RealTilePipeLaunchers<F64_16, 64, 2, double>::make("F64_16_rtpipe64", 40), RealTilePipeLaunchers<T64_32, 32, 1, double>::make("T64_32_rtpipe32", 40),
        V(F64_4, 2, 10), V(F64_8, 2, 10), V(F64_16, 2, 10), V(F64_32, 2, 10), V(F64_64, 2, 10), V(F64_128, 2, 10),
        V(F64_256, 2, 10), V(F64_512, 2, 10), V(F64_1024, 2, 10), V(F64_2048, 2, 10), V(F64_4096, 1, 10), V(F64_8192, 1, 10),
    };
    return v;
}
}  // namespace wfb
