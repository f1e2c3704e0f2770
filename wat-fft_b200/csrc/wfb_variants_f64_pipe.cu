// persistent TMA-pipelined f64 kernels: c2c interleaved and r2c / c2r
#include "wfb_registry.h"
namespace wfb {
#define VP(PL, X, MINB, PRIO) PipeLaunchers<double, PL, X, MINB>::make(#PL "_pipe" #X, PRIO)
#define VR(PL, X, MINB, ...) RealPipeLaunchers<double, PL, X, MINB>::make(#PL "_rpipe" #X, __VA_ARGS__)
#define VTS(PL, X, MINB, PRIO) PipeLaunchers<double, PL, X, MINB, 0, PADQ, true>::make(#PL "_pipe" #X "_ts", PRIO)
#define VRTS(PL, X, MINB, ...) RealPipeLaunchers<double, PL, X, MINB, 0, PADQ, true>::make(#PL "_rpipe" #X "_ts", __VA_ARGS__)
// + the last pass's twiddles held in registers across tiles (HT, see hoist_load in wfb_kernels.cuh)
#define VTSH(PL, X, MINB, PRIO) PipeLaunchers<double, PL, X, MINB, 0, PADQ, true, true>::make(#PL "_pipe" #X "_ts_ht", PRIO)
#define VRTSH(PL, X, MINB, ...) RealPipeLaunchers<double, PL, X, MINB, 0, PADQ, true, X, true>::make(#PL "_rpipe" #X "_ts_ht", __VA_ARGS__)
// + a higher minimum of resident CTAs (the register cap that goes with it): the f64 real kernels run 8 warps per SM
#define VRTSM(PL, X, MINB, ...) RealPipeLaunchers<double, PL, X, MINB, 0, PADQ, true>::make(#PL "_rpipe" #X "_ts_m" #MINB, __VA_ARGS__)
#define VRTSHM(PL, X, MINB, ...) RealPipeLaunchers<double, PL, X, MINB, 0, PADQ, true, X, true>::make(#PL "_rpipe" #X "_ts_ht_m" #MINB, __VA_ARGS__)
const std::vector<Variant> &variants_f64_pipe() {
    // priorities from the sweeps in profiles/ (c2c N = 256: the direct kernel wins; real transforms: the pipelined
    // kernels win up to N = 2048 since the Hermitian step moves only half its data through shared memory)
    static const std::vector<Variant> v = {
        VP(F64_256, 8, 2, 5), VP(F64_512, 4, 2, 30), VP(F64_1024, 2, 2, 30), VP(F64_2048, 1, 2, 30), VP(F64_4096, 1, 1, 30),
        // results leave as bulk stores out of the stage buffer (see k_c2c_pipe / k_real_pipe, TS): defaults from N = 256
        VTS(F64_64, 16, 2, 60), VTS(F64_128, 8, 2, 60), VRTS(F64_64, 16, 2, 60),
        VTS(F64_256, 8, 2, 60), VTS(F64_512, 4, 2, 60), VTS(F64_1024, 2, 2, 60), VTS(F64_2048, 1, 2, 60), VTS(F64_4096, 1, 1, 60),
        VRTS(F64_128, 16, 2, 60), VRTS(F64_256, 8, 2, 60), VRTS(F64_512, 4, 2, 60), VRTS(F64_1024, 2, 2, 60), VRTS(F64_2048, 2, 1, 60),
        // 16 KB tiles (f64 rows are twice as wide): +1..8 % at N >= 1024; single f64 rows are 16-byte multiples, no shift needed
        VTS(F64_512, 2, 2, 61), VTS(F64_1024, 1, 2, 61),
        VRTS(F64_512, 2, 2, 61), VRTS(F64_1024, 1, 2, 61), VRTS(F64_2048, 1, 1, 61),
        // one exchange instead of two at M = 512 (r2c f64 N = 1024: 79 -> 92 %, c2r: 93 -> 104 %); the radix-4 cores (M = 1024,
        // 4096) and M = 2048 do not factor into two passes of <= 32 values per thread
        VRTS(D32_512, 2, 1, 62),
        // hoisted last-pass twiddles where the L1 data pipe is the bound (ncu r02: 84 % at r2c N = 2048, 78 % at 4096, 71 % at c2c 4096)
        // Measured (profiles/r02_sweep.md, burst clocks, fraction of the measured HBM peak, plain -> hoisted):
        //   c2c N = 4096: 0.86 -> 0.875;  c2c N = 1024, 2048: 2-3 % SLOWER (not pipe-bound there)
        //   r2c N = 2048: 0.81 -> 0.83, with 6 CTAs/SM (168-register cap) 0.87;  c2r N = 2048: 0.85 -> 0.90 -> 0.96
        //   r2c N = 4096: 0.76 -> 0.78 (3 CTAs/SM: 0.78);  c2r N = 4096: 0.81 -> 0.79 -> 0.82;  r2c N = 1024 (32 values per thread): slower (spills)
        VTSH(F64_1024, 1, 2, 40), VTSH(F64_2048, 1, 2, 40), VTSH(F64_4096, 1, 1, 62),
        VRTSH(D32_512, 2, 1, 40), VRTSH(F64_1024, 1, 2, 40), VRTSH(F64_2048, 1, 1, 40),
        VRTSM(F64_1024, 1, 6, 39), VRTSHM(F64_1024, 1, 6, 63), VRTSM(F64_2048, 1, 3, 39), VRTSHM(F64_2048, 1, 3, 63), VRTSM(F64_512, 2, 8, 39),
        VR(F64_128, 16, 2, 30), VR(F64_256, 8, 2, 30), VR(F64_512, 4, 2, 5, 30), VR(F64_1024, 2, 2, 30), VR(F64_2048, 2, 1, 5),
    };
    return v;
}
}  // namespace wfb
