// hp_decode_staged.cuh - the production shape of accuracy() for 64x64 maps on sm_100a
// (utils/keypoint_detection.py:63-92: decode output and target, PCK counts, per-joint accuracies).
//
// Same shape as hp_loss_staged.cuh: persistent blocks, every warp owns whole map pairs and a private 32 KB
// shared-memory stage (output map + target map) filled by the copy engine; the argmax of both maps runs on LDS
// traffic with the first-index rules of the headline kernel (per-lane first iteration of the lane's maximum,
// redux.sync max over the warp, smallest (iteration, lane) key among the holders, first matching component);
// a map whose element sum is NaN (a NaN, or +inf with -inf) is rescanned from the stage with numpy's exact rules
// before the stage is re-filled.  PCK counters are per-block shared-memory integers, flushed once per block.
// Algorithmic bytes per map: 2*H*W*4 read + 8 written.  Roofline: HBM.
#pragma once
#include <cstdlib>

#include "hp_common.cuh"
#include "hp_decode.cuh"
#include "hp_tma.cuh"

namespace hp {

__device__ __forceinline__ ArgMax staged_warp_argmax(ArgMax am, int lane) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        ArgMax b;
        b.v = __shfl_xor_sync(0xffffffffu, am.v, o);
        b.i = __shfl_xor_sync(0xffffffffu, am.i, o);
        am = ((lane & o) == 0) ? am_merge(am, b) : am_merge(b, am);
    }
    return am;
}

// argmax over this warp's part of a map in shared memory: float4 number f0 + i*STEP, i < NITW, of the map (element
// e = component e & 3 of float4 number e >> 2; f0 = 32*h + lane, so a warp reads 512 consecutive bytes per iteration).
// numpy semantics (first index among equals, NaN wins); every lane returns the result.
template <int NITW, int STEP>
__device__ __forceinline__ ArgMax staged_argmax(const float4* __restrict__ buf, int f0, int lane) {
    float run = -INFINITY;
    int best_i = 0;
    float2 witness = make_float2(0.f, 0.f);
#pragma unroll 8
    for (int i = 0; i < NITW; ++i) {
        const float4 v = buf[f0 + i * STEP];
        const float t = fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w));
        best_i = (t > run) ? i : best_i;  // strict: the earlier iteration keeps ties
        run = fmaxf(run, t);
        witness = __fadd2_rn(witness, __fadd2_rn(make_float2(v.x, v.y), make_float2(v.z, v.w)));
    }
    const float w = witness.x + witness.y;
    if (__any_sync(0xffffffffu, w != w)) {  // rare: a NaN (or inf - inf) somewhere in this part of the map
        ArgMax sx = am_init();
        for (int i = 0; i < NITW; ++i) am_scan4<true>(sx, buf[f0 + i * STEP], (f0 + i * STEP) * 4);
        return staged_warp_argmax(sx, lane);
    }
    const float cm = warp_max_f32(run);
    // first flat index of cm: the smallest float4 number among the lanes that hold it, then the component
    const unsigned key = (run == cm) ? static_cast<unsigned>(f0 + best_i * STEP) : 0x7fffffffu;
    const unsigned kmin = __reduce_min_sync(0xffffffffu, key);
    const float4 v = buf[kmin];
    const int comp = (v.x == cm) ? 0 : ((v.y == cm) ? 1 : ((v.z == cm) ? 2 : 3));
    return ArgMax{cm, static_cast<int>(kmin) * 4 + comp};
}

// fp32 pre-test of the PCK decision (the float64 arithmetic of the reference only runs within 1e-4 of the threshold;
// coordinates are small integers, so the fp32 distance carries ~3e-7 relative error)
struct PckPretest {
    float thr2_lo, thr2_hi, inv_nx, inv_ny;
};
inline PckPretest pck_pretest(int H, int W, double thr) {
    PckPretest p;
    const double t2 = thr * thr;
    p.thr2_lo = static_cast<float>(t2 * (1.0 - 1e-4));
    p.thr2_hi = static_cast<float>(t2 * (1.0 + 1e-4));
    if (thr <= 0.0) p.thr2_lo = p.thr2_hi = -1.0f;  // d < thr never holds
    p.inv_nx = static_cast<float>(10.0 / H);  // norm = (H/10, W/10) applied to (x, y)  (keypoint_detection.py:77)
    p.inv_ny = static_cast<float>(10.0 / W);
    return p;
}
static __device__ __noinline__ int pck_hit_exact(float px, float py, float tx, float ty, int H, int W, double thr) {
    int valid, hit;
    pck_one(px, py, tx, ty, H, W, thr, valid, hit);
    return hit;
}
// utils/keypoint_detection.py:26-34 in integers: idx < 2^24, so x = idx % W and y = floor(idx / W) are exact either way
__device__ __forceinline__ void staged_decode_xy(ArgMax a, int W, float& px, float& py) {
    const int qy = a.i / W, qx = a.i - qy * W;
    const float keep = (a.v > 0.0f) ? 1.0f : 0.0f;  // NaN -> 0
    px = static_cast<float>(qx) * keep;
    py = static_cast<float>(qy) * keep;
}

// NS stages (= map pairs in flight) per block, WPS warps share a stage (WPS = 2: even / odd iterations, the partial
// argmax pairs meet at one named barrier per map and merge with numpy's tie rules), BPS blocks per SM
template <int NIT, int NS, int WPS, int BPS>
__global__ void __launch_bounds__(32 * NS * WPS, BPS)
    accuracy_staged_kernel(const float* __restrict__ output, const float* __restrict__ target, int n_maps, int K, int H, int Wd,
                           double thr, const PckPretest pre, float* __restrict__ pred_xy, int32_t* __restrict__ counts_out,
                           double* __restrict__ acc_out, Workspace* __restrict__ ws) {
    extern __shared__ __align__(128) unsigned char s_dyn[];
    __shared__ int s_counts[2 * HP_MAX_K];
    __shared__ __align__(8) unsigned long long s_bar[NS];
    __shared__ ArgMax s_x[NS][2][2];  // partner's (output, target) partial results, double-buffered by the map's parity
    constexpr int HW = NIT * 128;
    constexpr uint32_t kMapBytes = static_cast<uint32_t>(HW) * 4u, kStageBytes = 2u * kMapBytes;
    static_assert((WPS == 1 || WPS == 2) && NS <= 4, "one warp or a pair per stage; pair_barrier covers 4 stages");
    constexpr int STEP = 32 * WPS, NITW = NIT / WPS;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int st = warp / WPS, h = warp % WPS;
    const bool leader = h == 0 && lane == 0;
    const int n_stages = static_cast<int>(gridDim.x) * NS, gs = static_cast<int>(blockIdx.x) * NS + st;
    const int n_mine = (n_maps > gs) ? (n_maps - gs + n_stages - 1) / n_stages : 0;
    unsigned char* stage = s_dyn + static_cast<size_t>(st) * kStageBytes;
    const uint32_t stage_u32 = smem_addr(stage), bar_u32 = smem_addr(&s_bar[st]);
    const uint64_t pol = l2_evict_first_policy();
    auto request = [&](int map) {
        mbar_arrive_expect_tx(bar_u32, kStageBytes);
        bulk_load(stage_u32, output + static_cast<size_t>(map) * HW, kMapBytes, bar_u32, pol);
        bulk_load(stage_u32 + kMapBytes, target + static_cast<size_t>(map) * HW, kMapBytes, bar_u32, pol);
    };
    if (leader) {
        mbar_init(bar_u32, 1);
        mbar_init_fence();
        if (n_mine > 0) request(gs);
    }
    for (int i = threadIdx.x; i < 2 * HP_MAX_K; i += 32 * NS * WPS) s_counts[i] = 0;
    __syncthreads();  // barriers and counters initialised

    const float4* o4 = reinterpret_cast<const float4*>(stage);
    const float4* t4 = o4 + HW / 4;
    const int f0 = 32 * h + lane;
    for (int jj = 0; jj < n_mine; ++jj) {
        const int map = gs + jj * n_stages;
        mbar_wait(bar_u32, static_cast<uint32_t>(jj) & 1u);
        ArgMax ao = staged_argmax<NITW, STEP>(o4, f0, lane);
        ArgMax at = staged_argmax<NITW, STEP>(t4, f0, lane);
        if (WPS == 2) {
            ArgMax* x = s_x[st][jj & 1];
            if (h == 1 && lane == 0) {
                x[0] = ao;
                x[1] = at;
            }
            pair_barrier(st);  // both warps have read the stage out; the partner's results are visible
            if (h == 0) {
                ao = am_merge(ao, x[0]);
                at = am_merge(at, x[1]);
            }
        } else {
            __syncwarp();  // both maps are read out
        }
        if (leader) {  // request the stage's next pair before closing this one
            if (jj + 1 < n_mine) request(map + n_stages);
            float px, py, tx, ty;
            staged_decode_xy(ao, Wd, px, py);
            staged_decode_xy(at, Wd, tx, ty);
            *reinterpret_cast<float2*>(pred_xy + 2 * static_cast<size_t>(map)) = make_float2(px, py);
            const int valid = (tx > 1.0f && ty > 1.0f) ? 1 : 0;  // keypoint_detection.py:44
            int hit = 0;
            if (valid) {
                const float da = (px - tx) * pre.inv_nx, db = (py - ty) * pre.inv_ny;
                const float d2 = fmaf(da, da, db * db);
                if (d2 < pre.thr2_lo) hit = 1;
                else if (!(d2 > pre.thr2_hi)) hit = pck_hit_exact(px, py, tx, ty, H, Wd, thr);
            }
            const int k = map % K;
            if (valid) atomicAdd(&s_counts[K + k], 1);
            if (hit) atomicAdd(&s_counts[k], 1);
        }
    }
    __syncthreads();  // the block's counters are final
    bool wrote = leader && n_mine > 0;
    for (int i = threadIdx.x; i < 2 * K; i += 32 * NS * WPS) {
        const int v = s_counts[i];
        if (v != 0) {
            atomicAdd(&ws->counts[i], v);
            wrote = true;
        }
    }
    if (last_block_arrives_writers(&ws->counter, gridDim.x, wrote)) pck_publish(ws, K, counts_out, acc_out);
}

template <int WPS>
static int launch_accuracy_staged_wps(const float* output, const float* target, int n_maps, int K, int H, int W, double thr,
                                      float* pred_xy, int32_t* counts, double* acc_out, Workspace* ws, cudaStream_t stream) {
    constexpr int NIT = 32, NS = 3, BPS = 2;
    static int sms_dev[64] = {};
    static bool attr_done_dev[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    int& sms = sms_dev[dev & 63];
    if (sms == 0) {
        sms = hp_device_sm_count();
        if (sms <= 0) sms = 148;
    }
    constexpr size_t smem = static_cast<size_t>(NS) * 2 * NIT * 512;
    bool& attr_done = attr_done_dev[dev & 63];
    if (!attr_done) {
        const cudaError_t e = cudaFuncSetAttribute(accuracy_staged_kernel<NIT, NS, WPS, BPS>,
                                                   cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) return fail(static_cast<int>(e), "hp_accuracy: %s", cudaGetErrorString(e));
        attr_done = true;
    }
    int grid = sms * BPS;
    const int need = (n_maps + NS - 1) / NS;
    if (grid > need) grid = need;
    accuracy_staged_kernel<NIT, NS, WPS, BPS><<<grid, 32 * NS * WPS, smem, stream>>>(
        output, target, n_maps, K, H, W, thr, pck_pretest(H, W, thr), pred_xy, counts, acc_out, ws);
    return launch_status("hp_accuracy");
}

// returns 1 when the shape is not covered (the caller then takes the block-per-map kernel), 0 when launched
static int launch_accuracy_staged(const float* output, const float* target, int n_maps, int K, int H, int W, double thr,
                                  float* pred_xy, int32_t* counts, double* acc_out, Workspace* ws, cudaStream_t stream) {
    if (H * W != 4096 || !aligned16(output) || !aligned16(target) || !aligned8(pred_xy)) return 1;
    // HP_ACC_SHAPE: 'b' = the block-per-map kernel, '1' / '2' = warps per stage (comparison runs, tests)
    int wps = 2;
    if (const char* e = getenv("HP_ACC_SHAPE")) {
        if (e[0] == 'b') return 1;
        if (e[0] == '1') wps = 1;
        if (e[0] == '2') wps = 2;
    }
    return wps == 2 ? launch_accuracy_staged_wps<2>(output, target, n_maps, K, H, W, thr, pred_xy, counts, acc_out, ws, stream)
                    : launch_accuracy_staged_wps<1>(output, target, n_maps, K, H, W, thr, pred_xy, counts, acc_out, ws, stream);
}

}  // namespace hp
