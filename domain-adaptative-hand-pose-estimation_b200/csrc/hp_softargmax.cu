// hp_softargmax.cu - soft-argmax decode.
//
// Replaces  compute_uv_from_heatmaps3   utils/keypoint_detection.py:209-239
//     softmax(100 * heatmap) over H*W, expectation of the row / column index, output (E[col], E[row]) * 4.
// The reference materialises the scaled map, the softmax, two coordinate grids and two products (seven full-size
// temporaries); here a block of 128 threads holds a map in registers (one HBM read), takes the maximum, then the three
// sums (sum e, sum e*col, sum e*row) against it.  Maps larger than one register tile are walked twice (the second
// pass is served by L2).  Roofline: HBM, H*W*4 + 8 bytes per map.
#include "hp_common.cuh"

namespace hp {

constexpr int kSATPM = 128;
constexpr int kSANV = 8;

__device__ __forceinline__ float block_reduce_max_128(float x, float* s_buf) {
    x = warp_max_f32(x);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) s_buf[warp] = x;
    __syncthreads();
    return fmaxf(fmaxf(s_buf[0], s_buf[1]), fmaxf(s_buf[2], s_buf[3]));
}
__device__ __forceinline__ float block_reduce_sum_128(float x, float* s_buf) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) s_buf[warp] = x;
    __syncthreads();
    return (s_buf[0] + s_buf[1]) + (s_buf[2] + s_buf[3]);
}

template <int MODE>
__global__ void __launch_bounds__(kSATPM)
    soft_argmax_kernel(const float* __restrict__ heat, int n_maps, int HW, FastDiv wdiv, float beta, float scale,
                       float* __restrict__ out_uv) {
    __shared__ float s_buf[4];
    const int t = threadIdx.x;
    const int ntiles = tiles_for<kSATPM, kSANV>(HW);
    for (int map = blockIdx.x; map < n_maps; map += gridDim.x) {
        const float* pm = heat + static_cast<size_t>(map) * HW;
        float4 v[kSANV];
        // ---- pass 1: maximum of beta * h (a NaN anywhere makes every sum NaN below, like the reference) -------------
        float lm = -INFINITY;
        for (int tile = 0; tile < ntiles; ++tile) {
            load_tile<kSATPM, kSANV, MODE>(pm, HW, tile, t, beta >= 0.0f ? -INFINITY : INFINITY, v);
#pragma unroll
            for (int j = 0; j < kSANV; ++j)
                lm = fmaxf(lm, fmaxf(fmaxf(beta * v[j].x, beta * v[j].y), fmaxf(beta * v[j].z, beta * v[j].w)));
        }
        const float M = block_reduce_max_128(lm, s_buf);
        const float ms = (M == -INFINITY) ? 0.0f : M;
        const float bl = beta * kLog2e, mb = -ms * kLog2e;
        // ---- pass 2: sum e, sum e * col, sum e * row with e = exp(beta * h - M) ---------------------------------------
        float s = 0.f, sc = 0.f, sr = 0.f;
        for (int tile = 0; tile < ntiles; ++tile) {
            if (ntiles > 1) load_tile<kSATPM, kSANV, MODE>(pm, HW, tile, t, beta >= 0.0f ? -INFINITY : INFINITY, v);
#pragma unroll
            for (int j = 0; j < kSANV; ++j) {
                const int idx0 = tile * (kSATPM * kSANV * 4) + (j * kSATPM + t) * 4;
                if (idx0 >= HW) continue;
                uint32_t row, col;
                wdiv.divmod(static_cast<uint32_t>(idx0), row, col);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    if (MODE != WALK_EXACT && idx0 + c >= HW) continue;
                    uint32_t r = row, cc = col + c;
                    if (MODE == WALK_SCALAR && cc >= wdiv.d) wdiv.divmod(static_cast<uint32_t>(idx0 + c), r, cc);
                    const float e = exp2f(fmaf(f4_get(v[j], c), bl, mb));
                    s += e;
                    sc = fmaf(e, static_cast<float>(cc), sc);
                    sr = fmaf(e, static_cast<float>(r), sr);
                }
            }
        }
        const float S = block_reduce_sum_128(s, s_buf);
        const float SC = block_reduce_sum_128(sc, s_buf);
        const float SR = block_reduce_sum_128(sr, s_buf);
        if (t == 0) {
            out_uv[2 * map + 0] = scale * __fdiv_rn(SC, S);
            out_uv[2 * map + 1] = scale * __fdiv_rn(SR, S);
        }
    }
}

}  // namespace hp

using namespace hp;

extern "C" HP_API int hp_soft_argmax(const float* heat, int n_maps, int H, int W, float beta, float scale, float* out_uv,
                                     hp_stream_t stream) {
    HP_REQUIRE(heat && out_uv, HP_ERR_NULL, "hp_soft_argmax: null pointer");
    HP_REQUIRE(n_maps >= 0 && H > 0 && W > 0 && static_cast<long long>(H) * W < (1ll << 28), HP_ERR_SHAPE,
               "hp_soft_argmax: bad shape n_maps=%d H=%d W=%d", n_maps, H, W);
    HP_REQUIRE(aligned4(heat), HP_ERR_ALIGN, "hp_soft_argmax: misaligned input");
    if (n_maps == 0) return HP_OK;
    const int HW = H * W;
    int sms = hp_device_sm_count();
    if (sms <= 0) sms = 148;
    const int grid = n_maps < sms * 16 ? n_maps : sms * 16;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const FastDiv wdiv(static_cast<uint32_t>(W));
    if (aligned16(heat) && W % 4 == 0) {
        if (HW == kSATPM * kSANV * 4) soft_argmax_kernel<WALK_EXACT><<<grid, kSATPM, 0, s>>>(heat, n_maps, HW, wdiv, beta, scale, out_uv);
        else soft_argmax_kernel<WALK_VEC><<<grid, kSATPM, 0, s>>>(heat, n_maps, HW, wdiv, beta, scale, out_uv);
    } else {
        soft_argmax_kernel<WALK_SCALAR><<<grid, kSATPM, 0, s>>>(heat, n_maps, HW, wdiv, beta, scale, out_uv);
    }
    return launch_status("hp_soft_argmax");
}
