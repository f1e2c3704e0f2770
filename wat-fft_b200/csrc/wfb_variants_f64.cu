// f64 kernels: c2c interleaved, r2c, c2r (extension)
#include "wfb_registry.h"
namespace wfb {
#define V(PL, MINB, PRIO) Launchers<double, PL, XROWS(PL::T), MINB, false>::make(#PL, PRIO)
const std::vector<Variant> &variants_f64() {
    static const std::vector<Variant> v = {
        V(F64_4, 2, 10), V(F64_8, 2, 10), V(F64_16, 2, 10), V(F64_32, 2, 10), V(F64_64, 2, 10), V(F64_128, 2, 10),
        V(F64_256, 2, 10), V(F64_512, 2, 10), V(F64_1024, 2, 10), V(F64_2048, 2, 10), V(F64_4096, 1, 10), V(F64_8192, 1, 10),
    };
    return v;
}
}  // namespace wfb
