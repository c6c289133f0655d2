// hp_target.cu - batched Gaussian target generation (a5).
//
// Replaces the per-sample numpy loop of uda/dataset/util.py:9-68 (generate_target).  Write-only
// kernel: each block owns MPB consecutive maps, computes their centres once (float64 index
// arithmetic, truncation toward zero, util.py:36-46) and streams 128-bit stores.
// Roofline: HBM; algorithmic bytes per map = H*W*4 written + 24 read.
#include "hp_common.cuh"

namespace hp {

constexpr int kTargetThreads = 256;
constexpr int kTargetMaxMPB = 16;

template <bool VEC>
__global__ void __launch_bounds__(kTargetThreads)
    gaussian_target_kernel(const double* __restrict__ joints, const float* __restrict__ vis, int n_maps, int H, int W,
                           FastDiv wdiv, double sx, double sy, int tmp, const float* __restrict__ tab, int mpb,
                           float* __restrict__ target, float* __restrict__ weight) {
    extern __shared__ float s_tab[];
    __shared__ Centre s_c[kTargetMaxMPB];
    load_table(s_tab, tab, tmp);
    const int map0 = blockIdx.x * mpb;
    if (threadIdx.x < mpb && map0 + threadIdx.x < n_maps) {
        const int map = map0 + threadIdx.x;
        float w;
        s_c[threadIdx.x] = target_centre(joints[2 * map], joints[2 * map + 1], vis[map], sx, sy, W, H, w);
        weight[map] = w;
    }
    __syncthreads();
    const int HW = H * W;
    const int nmaps_here = min(mpb, n_maps - map0);
    float* out = target + static_cast<size_t>(map0) * HW;
    if (VEC) {
        const int hw4 = HW >> 2, total4 = nmaps_here * hw4;
        for (int v = threadIdx.x; v < total4; v += kTargetThreads) {
            const int g = v / hw4, e = (v - g * hw4) << 2;
            uint32_t y, x0;
            wdiv.divmod(static_cast<uint32_t>(e), y, x0);
            stg_stream4(reinterpret_cast<float4*>(out) + v, patch_at4(s_tab, tmp, s_c[g], static_cast<int>(x0), static_cast<int>(y)));
        }
    } else {
        const int total = nmaps_here * HW;
        for (int v = threadIdx.x; v < total; v += kTargetThreads) {
            const int g = v / HW, e = v - g * HW;
            uint32_t y, x;
            wdiv.divmod(static_cast<uint32_t>(e), y, x);
            out[v] = patch_at(s_tab, tmp, s_c[g], static_cast<int>(x), static_cast<int>(y));
        }
    }
}

}  // namespace hp

using namespace hp;

extern "C" HP_API int hp_gaussian_target(const double* joints, const float* vis, int n_maps, int H, int W,
                                         double stride_x, double stride_y, int tmp, const float* tab, float* target,
                                         float* weight, hp_stream_t stream) {
    HP_REQUIRE(joints && vis && tab && target && weight, HP_ERR_NULL, "hp_gaussian_target: null pointer");
    HP_REQUIRE(n_maps >= 0 && H > 0 && W > 0 && static_cast<long long>(H) * W < (1ll << 30), HP_ERR_SHAPE,
               "hp_gaussian_target: bad shape n_maps=%d H=%d W=%d", n_maps, H, W);
    HP_REQUIRE(tmp >= 0 && tmp <= 64 && stride_x > 0.0 && stride_y > 0.0, HP_ERR_ARG,
               "hp_gaussian_target: bad tmp=%d or stride", tmp);
    HP_REQUIRE(aligned8(joints) && aligned4(vis) && aligned4(target), HP_ERR_ALIGN, "hp_gaussian_target: misaligned");
    if (n_maps == 0) return HP_OK;
    const int HW = H * W;
    int mpb = (kTargetThreads * 16) / HW;  // aim at >= 16 floats written per thread
    mpb = mpb < 1 ? 1 : (mpb > kTargetMaxMPB ? kTargetMaxMPB : mpb);
    const int grid = (n_maps + mpb - 1) / mpb;
    const bool vec = (W % 4 == 0) && aligned16(target);
    const size_t smem = table_bytes(tmp);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (vec)
        gaussian_target_kernel<true><<<grid, kTargetThreads, smem, s>>>(joints, vis, n_maps, H, W, FastDiv(W), stride_x,
                                                                       stride_y, tmp, tab, mpb, target, weight);
    else
        gaussian_target_kernel<false><<<grid, kTargetThreads, smem, s>>>(joints, vis, n_maps, H, W, FastDiv(W),
                                                                        stride_x, stride_y, tmp, tab, mpb, target, weight);
    return launch_status("hp_gaussian_target");
}
