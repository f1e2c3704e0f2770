// persistent TMA-pipelined f64 kernels: c2c interleaved and r2c / c2r
#include "wfb_registry.h"
namespace wfb {
#define VP(PL, X, MINB, PRIO) PipeLaunchers<double, PL, X, MINB>::make(#PL "_pipe" #X, PRIO)
#define VR(PL, X, MINB, PRIO) RealPipeLaunchers<double, PL, X, MINB>::make(#PL "_rpipe" #X, PRIO)
const std::vector<Variant> &variants_f64_pipe() {
    // priorities from gpurun_out/sweep20 (after the FMA-fused butterfly the FP64-pipe-bound direct kernels
    // beat the pipelined ones for the real transforms and at N = 256)
    static const std::vector<Variant> v = {
        VP(F64_256, 8, 2, 5), VP(F64_512, 4, 2, 30), VP(F64_1024, 2, 2, 30), VP(F64_2048, 1, 2, 30), VP(F64_4096, 1, 1, 30),
        VR(F64_128, 16, 2, 5), VR(F64_256, 8, 2, 5), VR(F64_512, 4, 2, 5), VR(F64_1024, 2, 2, 5), VR(F64_2048, 2, 1, 5),
    };
    return v;
}
}  // namespace wfb
