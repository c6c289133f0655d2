// hp_pipeline_parts.cuh - small pieces shared by the shapes of the fused gen+loss+decode+PCK kernel: the per-lane
// patch table entry, the per-warp exact loss accumulators, the 3-value transpose reduction.
// (Until round 2 these lived next to the register-tile kernels of round 1 - pipeline_tiles*, 28.4 us per launch - which
// the TMA-staged kernel superseded and which have been retired; the history is in profiles/r1_pipeline_history.md.)
#pragma once
#include "hp_common.cuh"
#include "hp_pipeline_common.cuh"

namespace hp {

constexpr int kTileMaxPatch = 6;    // patch pixels per lane of the closing warp: (2*tmp+1)^2 <= 192

// per-lane patch slots: offset from the centre and the target terms that do not depend on the prediction
struct PatchSlot {
    int dx, dy;        // dx = 1<<20 for unused slots (never in bounds)
    float t, ulogu;    // target value, (t+eps)*ln(t+eps)
};

// Sum three values over the warp with 6 shuffles instead of 15: after each exchange a lane keeps half of
// its values.  On return lanes 0-7 hold sum(a), lanes 8-15 sum(b), lanes 16-23 sum(c), lanes 24-31 zero.
__device__ __forceinline__ float warp_sum3_scattered(float a, float b, float c, int lane) {
    const bool hi16 = (lane & 16) != 0;
    float k0 = hi16 ? c : a, k1 = hi16 ? 0.0f : b;  // low half keeps (a, b), high half keeps (c, 0)
    const float s0 = hi16 ? a : c, s1 = hi16 ? b : 0.0f;
    k0 += __shfl_xor_sync(0xffffffffu, s0, 16);
    k1 += __shfl_xor_sync(0xffffffffu, s1, 16);
    const bool hi8 = (lane & 8) != 0;
    float k = hi8 ? k1 : k0;
    const float s = hi8 ? k0 : k1;
    k += __shfl_xor_sync(0xffffffffu, s, 8);
    k += __shfl_xor_sync(0xffffffffu, k, 4);
    k += __shfl_xor_sync(0xffffffffu, k, 2);
    k += __shfl_xor_sync(0xffffffffu, k, 1);
    return k;
}

// per-warp exact loss accumulators (shared memory, touched by lane 0 of the owning warp only)
struct WarpLoss {
    long long fx[2];
    int cls[6];
};
static __device__ __noinline__ void warp_loss_add_nonfinite(WarpLoss* w, int which, double v) {
    if (v != v) w->cls[3 * which + 0] += 1;
    else if (v >= kFxLimit) w->cls[3 * which + 1] += 1;
    else w->cls[3 * which + 2] += 1;
}
__device__ __forceinline__ void warp_loss_add(WarpLoss* w, int which, double v) {
    if (fabs(v) < kFxLimit) w->fx[which] += __double2ll_rn(v * 1099511627776.0);  // * 2^40, exact scaling
    else warp_loss_add_nonfinite(w, which, v);
}

}  // namespace hp
