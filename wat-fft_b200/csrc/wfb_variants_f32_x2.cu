// packed-FP32 (f32x2: FFMA2/FADD2/FMUL2) kernels, two batch rows per thread group
#include "wfb_registry.h"
namespace wfb {
#define V(PL, MINB, ...) Launchers<f32x2, PL, XROWS(PL::T), MINB, true>::make(#PL "_x2", __VA_ARGS__)
const std::vector<Variant> &variants_f32_x2() {
    static const std::vector<Variant> v = {
        // interleaved layouts: the direct packed kernels load one LDG.64 per value and win at N = 128, 256 (sweep_final)
        V(F32_128, 2, 20, -1, 40), V(F32_256, 2, 20, -1, 40), V(F32_512, 2, 20), V(F32_1024, 2, 20),
        // N >= 2048 removed: never within 10 % of the pipelined scalar kernels (profiles/r01_sweep.md)
    };
    return v;
}
}  // namespace wfb
