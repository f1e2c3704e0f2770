// scalar-lane f32 kernels with direct global loads/stores: c2c (split + interleaved), r2c, c2r
#include "wfb_registry.h"
namespace wfb {
#define V(PL, MINB, PRIO) Launchers<float, PL, XROWS(PL::T), MINB, true>::make(#PL, PRIO)
const std::vector<Variant> &variants_f32_direct() {
    static const std::vector<Variant> v = {
        V(F32_4, 2, 10), V(F32_8, 2, 10), V(F32_16, 2, 10), V(F32_32, 2, 10), V(F32_64, 2, 10), V(F32_128, 2, 10),
        V(F32_256, 2, 10), V(F32_512, 2, 10), V(F32_1024, 2, 10), V(F32_2048, 2, 10), V(F32_4096, 2, 10), V(F32_8192, 1, 10),
    };
    return v;
}
}  // namespace wfb
