// persistent TMA-pipelined c2c kernels
#include "wfb_registry.h"
namespace wfb {
#define VP(PL, X, MINB, PRIO) PipeLaunchers<float, PL, X, MINB>::make(#PL "_pipe" #X, PRIO)
#define VPR(PL, X, MINB, PRIO) PipeLaunchers<float, PL, X, MINB, 1>::make(#PL "_pipe" #X "_rc", PRIO)
#define VP64(PL, X, MINB, PRIO) PipeLaunchers<float, PL, X, MINB, 0, 64>::make(#PL "_pipe" #X, PRIO)
#define VP32(PL, X, MINB, PRIO) PipeLaunchers<float, PL, X, MINB, 0, 32>::make(#PL "_pipe" #X, PRIO)
#define VTS(PL, X, MINB, RC, PQ, PRIO) PipeLaunchers<float, PL, X, MINB, RC, PQ, true>::make(#PL "_pipe" #X "_ts", PRIO)
// row copies in groups of RG rows per bulk copy (see RC in k_c2c_pipe)
#define VTSG(PL, X, MINB, RG, PQ, ...) PipeLaunchers<float, PL, X, MINB, RG, PQ, true>::make(#PL "_pipe" #X "_ts_g" #RG, __VA_ARGS__)
#define VTSM(PL, X, MINB, RC, PQ, PRIO) PipeLaunchers<float, PL, X, MINB, RC, PQ, true>::make(#PL "_pipe" #X "_ts_m" #MINB, PRIO)
#define VRP(PL, MINB, TS, NAME, PRIO) RegPipeLaunchers<float, PL, MINB, TS>::make(#PL NAME, PRIO)
#define VP2(PL, X, MINB, PRIO) PipeLaunchers<f32x2, PL, X, MINB>::make(#PL "_pipe" #X "_x2", PRIO)
const std::vector<Variant> &variants_f32_pipe() {
    // 16 KB tiles are the sweet spot for the dynamically scheduled pipeline (8 KB and 32 KB tiles lose 10-30 %).
    // priorities from the SUSTAINED sweep (tools/sweep.py --sustain 0.6, power-capped clocks, which is what
    // bench.py's long step sees; the burst ranking differs at N = 256 and 1024): higher = default
    static const std::vector<Variant> v = {
        VPR(F32_128, 16, 2, 34),
        VP(F32_256, 8, 2, 36), VP(F32_256, 16, 2, 30), VP2(F32_256, 8, 2, 33),
        VP(F32_512, 4, 2, 34), VP2(F32_512, 4, 2, 30),
        // 32 values per thread, one exchange: equal at burst clocks, 7-11 % faster power-capped (fewer instructions)
        VP(P32_512, 4, 2, 35), VP32(P32_1024, 2, 2, 34),
        VP(F32_1024, 2, 2, 33),
        VP(F32_2048, 1, 4, 30),
        VP(F32_4096, 1, 2, 30), VP(F32_8192, 1, 1, 30),
        // results leave through the stage buffer as bulk stores instead of per-thread STG
        // (profiles/: +1..4 % at burst clocks, +3..8 % power-capped; N = 4096: 86 -> 92 % at burst, equal power-capped)
        // N = 128: row copies in groups of 2 rows (32 bulk copies per tile and direction instead of 64): 177 -> 156 M executed
        // instructions per launch, power-capped split 0.92 / 0.91 -> 1.00 / 1.00 (fwd / inv), burst and interleaved unchanged
        // (1.03); groups of 4 (145 M instructions): 1.00 / 0.99
        VTSG(F32_128, 16, 2, 2, 16, 61), VTS(F32_128, 16, 2, 1, 16, 60), VTSG(F32_128, 16, 2, 4, 16, 59),
        // N = 256, 512 (T = 16): on the split planes a warp reads two rows side by side with 32-bit accesses, and in a
        // dense tile both sit on the same banks (ncu: 41 % of the shared-memory wavefronts replayed at both sizes).
        // Grouped copies put them 16 banks apart: power-capped 512 0.996 -> 1.016, 256 1.002 -> 1.004 on a fast box
        // (HBM-bound there); burst 1.040 -> 1.040 / 1.030 -> 1.023.  Defaults for the split layout; the interleaved
        // layout (64-bit accesses: one row per half-warp, no conflict) keeps the dense tile.
        VTSG(F32_256, 8, 2, 4, 16, 61, -1, 58), VTSG(P32_512, 4, 2, 2, 16, 61, -1, 58), VTS(F32_256, 8, 2, 0, 16, 60), VTS(P32_512, 4, 2, 0, 16, 60), VTS(P32_1024, 2, 2, 0, 32, 60),
        VTS(F32_2048, 1, 4, 0, 16, 60), VTS(P64_4096, 1, 1, 0, 64, 20), VTS(F32_4096, 1, 2, 0, 16, 60), VTS(F32_8192, 1, 1, 0, 16, 60),
        // N = 8192: three passes (32 values per thread) instead of four: 69 -> 82 %
        VTS(P32_8192, 1, 1, 0, 16, 61),
        // three resident CTAs per SM (85-register cap) for the barrier-heavy three-pass plans
        VTSM(F32_4096, 1, 3, 0, 16, 19), VTSM(F32_2048, 1, 6, 0, 16, 19),
        // register-prefetch persistent kernels (k_c2c_rpf): the input never touches shared memory
        VRP(F32_4096, 2, true, "_rpf_ts", 18), VRP(F32_4096, 2, false, "_rpf", 18), VRP(F32_2048, 3, true, "_rpf_ts", 18), VRP(F32_2048, 4, false, "_rpf", 18),
        VP64(P64_4096, 1, 1, 31),   // one exchange: 3 % slower than F32_4096_pipe1 at burst clocks, 4 % faster power-capped
         VP64(P64_2048, 1, 1, 20), VP64(P64_1024, 2, 1, 20),
    };
    return v;
}
}  // namespace wfb
