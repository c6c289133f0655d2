// hp_extras.cu - small boundary operators around the decode:
//   hp_argmax_decode_f64  get_max_preds for float64 heatmaps (utils/keypoint_detection.py:7-35 accepts any ndarray
//                         dtype; a cast to float32 could merge two distinct float64 maxima into a tie and move the argmax)
//   hp_refine_quarter     a13: the opt-in quarter-pixel refinement the north star names (NOT in the reference: default off)
//   hp_group_accuracy     keypoint_dataset.py:58-71 on the device: per-group means of the per-joint accuracies
// None of them is on the throughput path (K..B*K values of work); they exist so that a caller never has to leave the
// device - or the C ABI - for them.
#include "hp_common.cuh"

namespace hp {

// one warp per map, lanes stride over the elements; numpy argmax semantics (first index among equals, the first NaN wins)
__global__ void __launch_bounds__(128) decode_f64_kernel(const double* __restrict__ heat, int n_maps, int HW, int W,
                                                         float* __restrict__ preds, double* __restrict__ maxvals) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= n_maps) return;
    const double* m = heat + static_cast<size_t>(warp) * HW;
    double bv = -INFINITY;
    int bi = 0x7fffffff;
    bool bnan = false;
    for (int e = lane; e < HW; e += 32) {
        const double v = m[e];
        const bool vnan = v != v;
        // per lane the indices increase, so a strict comparison keeps the first among equals, and the first NaN stays
        const bool take = (bi == 0x7fffffff) || (!bnan && (vnan || v > bv));
        if (take) {
            bv = v;
            bi = e;
            bnan = vnan;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        const bool onan = ov != ov;
        bool take;
        if (oi == 0x7fffffff) take = false;
        else if (bi == 0x7fffffff) take = true;
        else if (bnan || onan) take = (bnan && onan) ? (oi < bi) : onan;
        else if (ov != bv) take = ov > bv;
        else take = oi < bi;
        if (take) {
            bv = ov;
            bi = oi;
            bnan = onan;
        }
    }
    if (lane == 0) {
        const float keep = (bv > 0.0) ? 1.0f : 0.0f;  // NaN -> 0 (keypoint_detection.py:31-34)
        // the reference converts the index to float32 BEFORE % and / (keypoint_detection.py:26-29): exact below 2^24
        const float fi = static_cast<float>(bi), fw = static_cast<float>(W);
        preds[2 * warp + 0] = fmodf(fi, fw) * keep;
        preds[2 * warp + 1] = floorf(fi / fw) * keep;
        maxvals[warp] = bv;
    }
}

// coord += 0.25 * sign(hm[y][x+1] - hm[y][x-1]) (resp. y) for maxima strictly inside the map (1 < x < W-1, 1 < y < H-1):
// the standard rule of heatmap pose estimators.  One thread per map, four loads.
__global__ void refine_quarter_kernel(const float* __restrict__ heat, int n_maps, int H, int W, float* __restrict__ preds) {
    const int map = blockIdx.x * blockDim.x + threadIdx.x;
    if (map >= n_maps) return;
    const float fx = preds[2 * map + 0], fy = preds[2 * map + 1];
    const int px = static_cast<int>(floorf(fx + 0.5f)), py = static_cast<int>(floorf(fy + 0.5f));
    if (px > 1 && px < W - 1 && py > 1 && py < H - 1) {
        const float* m = heat + static_cast<size_t>(map) * H * W;
        const float dx = m[py * W + px + 1] - m[py * W + px - 1];
        const float dy = m[(py + 1) * W + px] - m[(py - 1) * W + px];
        const float sx = (dx > 0.0f) ? 1.0f : ((dx < 0.0f) ? -1.0f : 0.0f);  // numpy sign: 0 at 0, NaN -> no move here
        const float sy = (dy > 0.0f) ? 1.0f : ((dy < 0.0f) ? -1.0f : 0.0f);
        preds[2 * map + 0] = fx + 0.25f * sx;
        preds[2 * map + 1] = fy + 0.25f * sy;
    }
}

// out[g] = sum(acc[index[j]] for j in [offsets[g], offsets[g+1])) / count, summed left to right from 0 like Python's
// sum() (keypoint_dataset.py:68-70) so the float64 result is bit-equal to the reference's
__global__ void group_accuracy_kernel(const double* __restrict__ acc, const int* __restrict__ offsets,
                                      const int* __restrict__ index, int G, double* __restrict__ out) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= G) return;
    double s = 0.0;
    const int b = offsets[g], e = offsets[g + 1];
    for (int j = b; j < e; ++j) s = __dadd_rn(s, acc[index[j]]);
    out[g] = __ddiv_rn(s, static_cast<double>(e - b));
}

}  // namespace hp

using namespace hp;

extern "C" HP_API int hp_argmax_decode_f64(const double* heat, int n_maps, int H, int W, float* preds, double* maxvals,
                                           hp_stream_t stream) {
    HP_REQUIRE(heat && preds && maxvals, HP_ERR_NULL, "hp_argmax_decode_f64: null pointer");
    HP_REQUIRE(n_maps >= 0 && H > 0 && W > 0 && static_cast<long long>(H) * W < (1ll << 24), HP_ERR_SHAPE,
               "hp_argmax_decode_f64: bad shape n_maps=%d H=%d W=%d", n_maps, H, W);
    HP_REQUIRE(aligned8(heat) && aligned8(maxvals) && aligned4(preds), HP_ERR_ALIGN, "hp_argmax_decode_f64: misaligned");
    if (n_maps == 0) return HP_OK;
    const int warps_per_block = 4;
    decode_f64_kernel<<<(n_maps + warps_per_block - 1) / warps_per_block, 32 * warps_per_block, 0,
                        static_cast<cudaStream_t>(stream)>>>(heat, n_maps, H * W, W, preds, maxvals);
    return launch_status("hp_argmax_decode_f64");
}

extern "C" HP_API int hp_refine_quarter(const float* heat, int n_maps, int H, int W, float* preds, hp_stream_t stream) {
    HP_REQUIRE(heat && preds, HP_ERR_NULL, "hp_refine_quarter: null pointer");
    HP_REQUIRE(n_maps >= 0 && H > 0 && W > 0, HP_ERR_SHAPE, "hp_refine_quarter: bad shape n_maps=%d H=%d W=%d", n_maps, H, W);
    if (n_maps == 0) return HP_OK;
    refine_quarter_kernel<<<(n_maps + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(heat, n_maps, H, W, preds);
    return launch_status("hp_refine_quarter");
}

extern "C" HP_API int hp_group_accuracy(const double* acc, const int32_t* offsets, const int32_t* index, int n_groups,
                                        double* out, hp_stream_t stream) {
    HP_REQUIRE(acc && offsets && index && out, HP_ERR_NULL, "hp_group_accuracy: null pointer");
    HP_REQUIRE(n_groups >= 0, HP_ERR_SHAPE, "hp_group_accuracy: n_groups=%d", n_groups);
    if (n_groups == 0) return HP_OK;
    group_accuracy_kernel<<<(n_groups + 63) / 64, 64, 0, static_cast<cudaStream_t>(stream)>>>(acc, offsets, index, n_groups, out);
    return launch_status("hp_group_accuracy");
}
