// hp_pipeline_stream.cuh - warp-per-map shape of the fused gen+loss+decode+PCK kernel.
//
// One WARP streams a whole map through double-buffered register tiles (8 x 128-bit loads per lane),
// no block barrier on the data path.  Used where a map is large or tiny relative to a block's share:
// 128x128 (16 tiles per map, many waves) and 16x16 (one quarter-size tile).  For 64x64 / 32x32 the
// TMA-staged shape (hp_pipeline_bulk.cuh) is faster: there a whole-map-per-warp work item is so long
// (~10 us) that the tail of a 2.3-wave grid idles the SMs 47 % of the time (profiles/r1_pipeline_v2.md).
// The instruction diet is the same: target-free hot loop (FMNMX3 max, ==max index scan, packed
// FFMA2/FADD2, MUFU.EX2), patch terms from an up-front re-read of the <=169 patch pixels.
#pragma once
#include "hp_pipeline_common.cuh"

namespace hp {

constexpr int kStreamWarps = 4;     // warps (= maps) per block
constexpr int kStreamMaxPatch = 6;  // patch pixels per lane: (2*tmp+1)^2 <= 192

template <int NV>
__device__ __forceinline__ void stream_load(const float4* __restrict__ m4, int tile, int lane, float4 (&v)[NV]) {
    const float4* p = m4 + tile * (32 * NV) + lane;
#pragma unroll
    for (int j = 0; j < NV; ++j) v[j] = ldg_stream4(p + j * 32);
}

struct StreamAcc {
    float best_v;
    int best_i;
    float m;      // running softmax max
    float2 s2;    // sum exp(p - m), two interleaved partial sums
    float2 sp2;   // sum p
    float2 spp2;  // sum p^2
};

template <int NV, int LOSS>
__device__ __forceinline__ void stream_tile(StreamAcc& A, const float4 (&v)[NV], int tile, int lane) {
    float tm = -INFINITY;
#pragma unroll
    for (int j = 0; j < NV; ++j) tm = fmaxf(tm, fmaxf(fmaxf(v[j].x, v[j].y), fmaxf(v[j].z, v[j].w)));
    int loc = 0;  // first position holding the tile maximum: scan downwards so the lowest survives
#pragma unroll
    for (int j = NV - 1; j >= 0; --j) {
        loc = (v[j].w == tm) ? (4 * j + 3) : loc;
        loc = (v[j].z == tm) ? (4 * j + 2) : loc;
        loc = (v[j].y == tm) ? (4 * j + 1) : loc;
        loc = (v[j].x == tm) ? (4 * j + 0) : loc;
    }
    if (tm > A.best_v) {  // strict: an earlier tile keeps ties
        A.best_v = tm;
        A.best_i = tile * (128 * NV) + (loc >> 2) * 128 + lane * 4 + (loc & 3);
    }
    if (LOSS & HP_LOSS_KL) {
        const float mn = fmaxf(A.m, tm);
        const float ms = (mn == -INFINITY) ? 0.0f : mn;
        const float scale = (A.m == -INFINITY) ? 0.0f : ex2_approx((A.m - ms) * kLog2e);
        A.s2.x *= scale;
        A.s2.y *= scale;
        A.m = mn;
        const float2 l2 = make_float2(kLog2e, kLog2e), mb2 = make_float2(-ms * kLog2e, -ms * kLog2e);
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            const float2 a0 = __ffma2_rn(make_float2(v[j].x, v[j].y), l2, mb2);
            const float2 a1 = __ffma2_rn(make_float2(v[j].z, v[j].w), l2, mb2);
            A.s2 = __fadd2_rn(A.s2, make_float2(ex2_approx(a0.x), ex2_approx(a0.y)));
            A.s2 = __fadd2_rn(A.s2, make_float2(ex2_approx(a1.x), ex2_approx(a1.y)));
        }
    }
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        const float2 lo = make_float2(v[j].x, v[j].y), hi = make_float2(v[j].z, v[j].w);
        A.sp2 = __fadd2_rn(A.sp2, __fadd2_rn(lo, hi));
        if (LOSS & HP_LOSS_MSE) {
            A.spp2 = __ffma2_rn(lo, lo, A.spp2);
            A.spp2 = __ffma2_rn(hi, hi, A.spp2);
        }
    }
}

template <int NV, int LOSS>
__global__ void __launch_bounds__(32 * kStreamWarps, 4) pipeline_stream_kernel(const PipeArgs a) {
    extern __shared__ float s_tab[];
    __shared__ double s_map[2][kStreamWarps];
    __shared__ BlockLoss s_loss;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int map = blockIdx.x * kStreamWarps + warp;
    const bool active = map < a.n_maps;
    const float* pm = a.pred + static_cast<size_t>(active ? map : 0) * a.HW;
    const float4* m4 = reinterpret_cast<const float4*>(pm);

    float4 buf0[NV], buf1[NV];
    if (active) stream_load<NV>(m4, 0, lane, buf0);  // in flight during the whole prologue
    load_table(s_tab, a.tab, a.tmp);
    if (threadIdx.x == 0) block_loss_zero(&s_loss);
    if (lane == 0) s_map[0][warp] = s_map[1][warp] = 0.0;
    __syncthreads();

    if (active) {
        float weight;
        const Centre c = pipe_centre(a, a.joints[2 * map], a.joints[2 * map + 1], a.vis[map], weight);
        const bool pasted = c.y != kNoPaste;
        PatchSums ps{0.f, 0.f, 0.f, 0.f, 0.f};
        {
            const int side = 2 * a.tmp + 1, n_patch = side * side;
            float pv[kStreamMaxPatch], tv[kStreamMaxPatch];
#pragma unroll
            for (int q = 0; q < kStreamMaxPatch; ++q) {
                const int i = lane + 32 * q;
                uint32_t ry, rx;
                a.sdiv.divmod(static_cast<uint32_t>(i), ry, rx);
                const int dx = static_cast<int>(rx) - a.tmp, dy = static_cast<int>(ry) - a.tmp;
                const int x = c.x + dx, y = c.y + dy;
                const bool in = pasted && i < n_patch && x >= 0 && x < a.W && y >= 0 && y < a.H;
                pv[q] = in ? ldg_stream1(pm + y * a.W + x) : 0.0f;
                tv[q] = in ? s_tab[dx * dx + dy * dy] : 0.0f;
            }
#pragma unroll
            for (int q = 0; q < kStreamMaxPatch; ++q) patch_pixel<LOSS>(ps, tv[q], pv[q], a.eps);
        }
        StreamAcc A;
        A.best_v = -INFINITY;
        A.best_i = 0;
        A.m = -INFINITY;
        A.s2 = A.sp2 = A.spp2 = make_float2(0.f, 0.f);
        for (int tile = 0; tile < a.ntiles; tile += 2) {
            if (tile + 1 < a.ntiles) stream_load<NV>(m4, tile + 1, lane, buf1);
            stream_tile<NV, LOSS>(A, buf0, tile, lane);
            if (tile + 1 < a.ntiles) {
                if (tile + 2 < a.ntiles) stream_load<NV>(m4, tile + 2, lane, buf0);
                stream_tile<NV, LOSS>(A, buf1, tile + 1, lane);
            }
        }
        ArgMax am = warp_argmax(ArgMax{A.best_v, A.best_i}, lane);
        float sum_exp = 0.f, sum_pp = 0.f;
        if (LOSS & HP_LOSS_KL) {
            const float ms = (am.v == -INFINITY) ? 0.0f : am.v;
            const float scale = (A.m == -INFINITY) ? 0.0f : ex2_approx((A.m - ms) * kLog2e);
            sum_exp = warp_sum((A.s2.x + A.s2.y) * scale);
            ps.up = warp_sum(ps.up);
            ps.ulogu = warp_sum(ps.ulogu);
            ps.u = warp_sum(ps.u);
            ps.p = warp_sum(ps.p);
        }
        const float sum_p = warp_sum(A.sp2.x + A.sp2.y);
        if (LOSS & HP_LOSS_MSE) {
            sum_pp = warp_sum(A.spp2.x + A.spp2.y);
            ps.e = warp_sum(ps.e);
        }
        if (sum_p != sum_p) {  // NaN (or +inf with -inf): exact numpy argmax rules, warp-uniform slow path
            ArgMax sx = am_init();
            for (int e4 = lane; e4 < a.HW / 4; e4 += 32) am_scan4<true>(sx, ldg_stream4(m4 + e4), e4 * 4);
            am = warp_argmax(sx, lane);
            sum_exp = __int_as_float(0x7fc00000);
        }
        uint32_t qy, qx;
        a.wdiv.divmod(static_cast<uint32_t>(am.i), qy, qx);
        const float keep = (am.v > 0.0f) ? 1.0f : 0.0f;
        const float px = static_cast<float>(qx) * keep, py = static_cast<float>(qy) * keep;
        const float tx = pasted ? static_cast<float>(c.x) : 0.0f, ty = pasted ? static_cast<float>(c.y) : 0.0f;
        int valid, hit;
        pipe_pck(a, px, py, tx, ty, valid, hit);
        double mse, kl;
        pipe_losses<LOSS>(a, c, weight, am.v, sum_exp, sum_p, sum_pp, ps, mse, kl);
        if (lane == 0) {
            a.pred_xy[2 * map + 0] = px;
            a.pred_xy[2 * map + 1] = py;
            if (a.maxvals) a.maxvals[map] = am.v;
            if (a.weight_out) a.weight_out[map] = weight;
            const int k = map % a.K;
            if (valid) atomicAdd(&a.ws->counts[a.K + k], 1);
            if (hit) atomicAdd(&a.ws->counts[k], 1);
            s_map[0][warp] = mse;
            s_map[1][warp] = kl;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int n_here = min(kStreamWarps, a.n_maps - static_cast<int>(blockIdx.x) * kStreamWarps);
        for (int w = 0; w < n_here; ++w) {
            if (LOSS & HP_LOSS_MSE) block_loss_add(&s_loss, 0, s_map[0][w]);
            if (LOSS & HP_LOSS_KL) block_loss_add(&s_loss, 1, s_map[1][w]);
        }
        block_loss_flush(&s_loss, a.ws);
    }
    if (pipeline_last_block(a.ws)) pipeline_publish(a);
}

}  // namespace hp
