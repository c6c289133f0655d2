// hp_softargmax.cu - soft-argmax decode.
//
// Replaces  compute_uv_from_heatmaps3   utils/keypoint_detection.py:209-239
//     softmax(100 * heatmap) over H*W, expectation of the row / column index, output (E[col], E[row]) * 4.
// The reference materialises the scaled map, the softmax, two coordinate grids and two products (seven full-size
// temporaries); here a block of 128 threads holds a map in registers (one HBM read), takes the maximum, then the three
// sums (sum e, sum e*col, sum e*row) against it.  Maps larger than one register tile are walked twice (the second
// pass is served by L2).  Roofline: HBM, H*W*4 + 8 bytes per map.
#include <cstdlib>

#include "hp_common.cuh"
#include "hp_tma.cuh"
#include "hp_pipeline_parts.cuh"  // warp_sum3_scattered

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

// ---------------------------------------------------------------------------------------------------------------------
// 4096-pixel maps (64x64): the shape of the fused pipeline kernel (hp_pipeline_bulk.cuh) - persistent blocks of 4 warps,
// 3 per SM; every warp owns one map at a time in a PRIVATE 16 KB shared-memory stage filled by the copy engine
// (cp.async.bulk -> mbarrier complete_tx) and re-armed by the warp that drained it; two passes over shared memory
// (maximum of beta * h, then sum e, sum e * col, sum e * row against it).  The block-per-map kernel above keeps HBM busy
// only through its own loads (0.59 of the roofline); here 192 KB per SM are requested whatever the warps are doing.
constexpr int kSAStageWarps = 4, kSAStageBlocks = 3;
__global__ void __launch_bounds__(32 * kSAStageWarps, kSAStageBlocks)
    soft_argmax_staged_kernel(const float* __restrict__ heat, int n_maps, FastDiv wdiv, float beta, float scale,
                              float* __restrict__ out_uv) {
    extern __shared__ __align__(128) unsigned char s_sa[];
    constexpr int W = kSAStageWarps, NITC = 32;
    constexpr uint32_t kBytes = NITC * 512;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_local = (n_maps - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
    const int n_mine = (n_local > warp) ? (n_local - warp + W - 1) / W : 0;
    unsigned char* my_stage = s_sa + static_cast<size_t>(warp) * kBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_sa + static_cast<size_t>(W) * kBytes) + warp;
    const uint32_t stage_u32 = smem_addr(my_stage), bar_u32 = smem_addr(bars);
    const size_t first_map = static_cast<size_t>(blockIdx.x) + static_cast<size_t>(warp) * gridDim.x;
    const size_t map_step = static_cast<size_t>(gridDim.x) * W;
    const uint64_t pol = l2_evict_first_policy();
    if (lane == 0) {
        mbar_init(bar_u32, 1);
        mbar_init_fence();
        if (n_mine > 0) {
            mbar_arrive_expect_tx(bar_u32, kBytes);
            bulk_load(stage_u32, heat + first_map * 4096, kBytes, bar_u32, pol);
        }
    }
    __syncwarp();
    // this lane's pixel coordinates: float4 number it * 32 + lane starts at flat index idx0 = 128 it + 4 lane
    const float bl = beta * kLog2e;
    const float4* buf = reinterpret_cast<const float4*>(my_stage);
    uint32_t parity = 0;
    for (int jj = 0; jj < n_mine; ++jj) {
        const size_t map = first_map + static_cast<size_t>(jj) * map_step;
        mbar_wait(bar_u32, parity);
        parity ^= 1u;
        // ---- pass 1: maximum of beta * h -------------------------------------------------------------------------------
        float lm = -INFINITY;
#pragma unroll 8
        for (int it = 0; it < NITC; ++it) {
            const float4 v = buf[it * 32 + lane];
            lm = fmaxf(lm, fmaxf(fmaxf(beta * v.x, beta * v.y), fmaxf(beta * v.z, beta * v.w)));
        }
        const float M = warp_max_f32(lm);
        const float ms = (M == -INFINITY) ? 0.0f : M;
        const float mb = -ms * kLog2e;
        // ---- pass 2: sum e, sum e * col, sum e * row with e = exp(beta * h - M) -----------------------------------------
        float s = 0.f, sc = 0.f, sr = 0.f;
#pragma unroll 4
        for (int it = 0; it < NITC; ++it) {
            const float4 v = buf[it * 32 + lane];
            uint32_t row, col;
            wdiv.divmod(static_cast<uint32_t>(128 * it + 4 * lane), row, col);
            const float fr = static_cast<float>(row), fc = static_cast<float>(col);
            const float e0 = exp2f(fmaf(v.x, bl, mb)), e1 = exp2f(fmaf(v.y, bl, mb));
            const float e2 = exp2f(fmaf(v.z, bl, mb)), e3 = exp2f(fmaf(v.w, bl, mb));
            const float es = (e0 + e1) + (e2 + e3);
            s += es;
            sr = fmaf(es, fr, sr);
            sc = fmaf(e0, fc, sc);
            sc = fmaf(e1, fc + 1.0f, sc);
            sc = fmaf(e2, fc + 2.0f, sc);
            sc = fmaf(e3, fc + 3.0f, sc);
        }
        __syncwarp();
        if (lane == 0 && jj + 1 < n_mine) {  // the stage has been read out: the warp's next map
            mbar_arrive_expect_tx(bar_u32, kBytes);
            bulk_load(stage_u32, heat + (map + map_step) * 4096, kBytes, bar_u32, pol);
        }
        const float r = warp_sum3_scattered(s, sc, sr, lane);
        const float S = __shfl_sync(0xffffffffu, r, 0), SC = __shfl_sync(0xffffffffu, r, 8), SR = __shfl_sync(0xffffffffu, r, 16);
        if (lane == 0) {
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
    // HP_SA_SHAPE=b keeps the block-per-map kernel at 64x64 (comparison runs)
    static const bool staged_on = []() {
        const char* e = getenv("HP_SA_SHAPE");
        return !(e && e[0] == 'b');
    }();
    if (aligned16(heat) && W % 4 == 0 && HW == 4096 && staged_on) {
        constexpr size_t smem = static_cast<size_t>(kSAStageWarps) * 32 * 512 + sizeof(uint64_t) * kSAStageWarps;
        static bool configured[16] = {};
        int dev = 0;
        cudaGetDevice(&dev);
        if (dev < 0 || dev >= 16 || !configured[dev]) {
            const cudaError_t e = cudaFuncSetAttribute(soft_argmax_staged_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
            if (e != cudaSuccess) return fail(static_cast<int>(e), "hp_soft_argmax: %s", cudaGetErrorString(e));
            if (dev >= 0 && dev < 16) configured[dev] = true;
        }
        const int g = n_maps < sms * kSAStageBlocks ? n_maps : sms * kSAStageBlocks;
        soft_argmax_staged_kernel<<<g, 32 * kSAStageWarps, smem, s>>>(heat, n_maps, wdiv, beta, scale, out_uv);
        return launch_status("hp_soft_argmax");
    }
    if (aligned16(heat) && W % 4 == 0) {
        if (HW == kSATPM * kSANV * 4) soft_argmax_kernel<WALK_EXACT><<<grid, kSATPM, 0, s>>>(heat, n_maps, HW, wdiv, beta, scale, out_uv);
        else soft_argmax_kernel<WALK_VEC><<<grid, kSATPM, 0, s>>>(heat, n_maps, HW, wdiv, beta, scale, out_uv);
    } else {
        soft_argmax_kernel<WALK_SCALAR><<<grid, kSATPM, 0, s>>>(heat, n_maps, HW, wdiv, beta, scale, out_uv);
    }
    return launch_status("hp_soft_argmax");
}
