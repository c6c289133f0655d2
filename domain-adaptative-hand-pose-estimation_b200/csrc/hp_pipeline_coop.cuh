// hp_pipeline_coop.cuh - the production shape of the fused gen+loss+decode+PCK kernel (64x64, 32x32).
//
// Persistent blocks of 4 warps; the 4 warps COOPERATE on one map (one 128*NV-element tile each),
// then move to the block's next map (static stride gridDim.x).  Lessons from the two earlier shapes
// (profiles/r1_pipeline_*.md):
//   * block-per-map with target math in the hot loop: 43 instructions/element, issue-bound (13 % HBM);
//   * warp-per-map streaming: 3x fewer instructions but a block lives ~10 us, the 2.3-wave grid leaves
//     the SMs idle 47 % of the time (24 % HBM).
// Here a work item is 1/9th of a block's share (~1.5 us), the next map's tile is already in flight in
// the second register buffer while the current one is reduced (16 KB per block always outstanding,
// 64 KB per SM), and the per-map serial work is spread over the warps:
//   warp (i mod 4)       re-reads the <=169 patch pixels (L2) and builds the target terms
//   every warp           hot loop over its tile: max, sum exp, sum p, sum p^2 (packed f32x2 math)
//   one barrier per map  per-warp partials meet in a double-buffered shared-memory slot
//   the OWNER warp       (lowest warp whose tile holds the map maximum) finds the first index of the
//                        maximum in its registers, closes the losses, decodes, scores PCK, publishes.
#pragma once
#include "hp_pipeline_common.cuh"

namespace hp {

constexpr int kCoopWarps = 4;
constexpr int kCoopMaxPatch = 6;  // patch pixels per lane of the patch warp: (2*tmp+1)^2 <= 192

struct CoopSlot {
    float wmax[kCoopWarps], s[kCoopWarps], sp[kCoopWarps], spp[kCoopWarps];
    PatchSums patch;
    int cx, cy;
    float weight;
};

template <int NV>
__device__ __forceinline__ void coop_load(const float* __restrict__ pred, size_t map, int HW, int warp, int lane,
                                          float4 (&v)[NV]) {
    const float4* p = reinterpret_cast<const float4*>(pred + map * HW) + warp * (32 * NV) + lane;
#pragma unroll
    for (int j = 0; j < NV; ++j) v[j] = ldg_stream4(p + j * 32);
}

template <int NV, int LOSS>
__device__ __forceinline__ void coop_map(const PipeArgs& a, const float4 (&v)[NV], int map, int iter, int warp, int lane,
                                         double jx, double jy, float jvis, const float* s_tab, CoopSlot* slot,
                                         BlockLoss* bl) {
    const float* pm = a.pred + static_cast<size_t>(map) * a.HW;
    // ---- patch warp: issue the patch-pixel loads first, they complete under the hot loop -------------
    const bool patch_warp = (iter & (kCoopWarps - 1)) == warp;
    float pv[kCoopMaxPatch], tv[kCoopMaxPatch];
    Centre c;
    float weight = 0.f;
    if (patch_warp) {
        c = pipe_centre(a, jx, jy, jvis, weight);
        const bool pasted = c.y != kNoPaste;
        const int side = 2 * a.tmp + 1, n_patch = side * side;
#pragma unroll
        for (int q = 0; q < kCoopMaxPatch; ++q) {
            const int i = lane + 32 * q;
            uint32_t ry, rx;
            a.sdiv.divmod(static_cast<uint32_t>(i), ry, rx);
            const int dx = static_cast<int>(rx) - a.tmp, dy = static_cast<int>(ry) - a.tmp;
            const int x = c.x + dx, y = c.y + dy;
            const bool in = pasted && i < n_patch && x >= 0 && x < a.W && y >= 0 && y < a.H;
            pv[q] = in ? ldg_stream1(pm + y * a.W + x) : 0.0f;
            tv[q] = in ? s_tab[dx * dx + dy * dy] : 0.0f;
        }
    }
    // ---- hot loop: this warp's tile -------------------------------------------------------------------
    float tm = -INFINITY;
#pragma unroll
    for (int j = 0; j < NV; ++j) tm = fmaxf(tm, fmaxf(fmaxf(v[j].x, v[j].y), fmaxf(v[j].z, v[j].w)));
    float2 s2 = make_float2(0.f, 0.f), sp2 = make_float2(0.f, 0.f), spp2 = make_float2(0.f, 0.f);
    if (LOSS & HP_LOSS_KL) {
        const float ms = (tm == -INFINITY) ? 0.0f : tm;
        const float2 l2 = make_float2(kLog2e, kLog2e), mb2 = make_float2(-ms * kLog2e, -ms * kLog2e);
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            const float2 a0 = __ffma2_rn(make_float2(v[j].x, v[j].y), l2, mb2);
            const float2 a1 = __ffma2_rn(make_float2(v[j].z, v[j].w), l2, mb2);
            s2 = __fadd2_rn(s2, make_float2(ex2_approx(a0.x), ex2_approx(a0.y)));
            s2 = __fadd2_rn(s2, make_float2(ex2_approx(a1.x), ex2_approx(a1.y)));
        }
    }
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        const float2 lo = make_float2(v[j].x, v[j].y), hi = make_float2(v[j].z, v[j].w);
        sp2 = __fadd2_rn(sp2, __fadd2_rn(lo, hi));
        if (LOSS & HP_LOSS_MSE) {
            spp2 = __ffma2_rn(lo, lo, spp2);
            spp2 = __ffma2_rn(hi, hi, spp2);
        }
    }
    // ---- warp reduction -> slot --------------------------------------------------------------------------
    const float wmax = warp_max(tm);
    float s = 0.f, spp = 0.f;
    if (LOSS & HP_LOSS_KL) {
        const float ws = (wmax == -INFINITY) ? 0.0f : wmax;
        const float scale = (tm == -INFINITY) ? 0.0f : ex2_approx((tm - ws) * kLog2e);
        s = warp_sum((s2.x + s2.y) * scale);
    }
    const float sp = warp_sum(sp2.x + sp2.y);
    if (LOSS & HP_LOSS_MSE) spp = warp_sum(spp2.x + spp2.y);
    PatchSums ps{0.f, 0.f, 0.f, 0.f, 0.f};
    if (patch_warp) {
#pragma unroll
        for (int q = 0; q < kCoopMaxPatch; ++q) patch_pixel<LOSS>(ps, tv[q], pv[q], a.eps);
        if (LOSS & HP_LOSS_KL) {
            ps.up = warp_sum(ps.up);
            ps.ulogu = warp_sum(ps.ulogu);
            ps.u = warp_sum(ps.u);
            ps.p = warp_sum(ps.p);
        }
        if (LOSS & HP_LOSS_MSE) ps.e = warp_sum(ps.e);
    }
    if (lane == 0) {
        slot->wmax[warp] = wmax;
        slot->s[warp] = s;
        slot->sp[warp] = sp;
        slot->spp[warp] = spp;
        if (patch_warp) {
            slot->patch = ps;
            slot->cx = c.x;
            slot->cy = c.y;
            slot->weight = weight;
        }
    }
    __syncthreads();
    // ---- who owns the maximum? ------------------------------------------------------------------------------
    float wm[kCoopWarps];
#pragma unroll
    for (int w = 0; w < kCoopWarps; ++w) wm[w] = slot->wmax[w];
    const float M = fmaxf(fmaxf(wm[0], wm[1]), fmaxf(wm[2], wm[3]));
    int owner = kCoopWarps - 1;
#pragma unroll
    for (int w = kCoopWarps - 2; w >= 0; --w) owner = (wm[w] == M) ? w : owner;
    if (warp != owner) return;

    // ---- owner: first index of M inside its tile (registers), then the per-map closure ----------------------
    int loc = 4 * NV;
#pragma unroll
    for (int j = NV - 1; j >= 0; --j) {
        loc = (v[j].w == M) ? (4 * j + 3) : loc;
        loc = (v[j].z == M) ? (4 * j + 2) : loc;
        loc = (v[j].y == M) ? (4 * j + 1) : loc;
        loc = (v[j].x == M) ? (4 * j + 0) : loc;
    }
    const int cand = (loc < 4 * NV) ? (warp * (128 * NV) + (loc >> 2) * 128 + lane * 4 + (loc & 3)) : 0x7fffffff;
    ArgMax am{M, warp_min_int(cand)};
    float sum_exp = 0.f, sum_p = 0.f, sum_pp = 0.f;
    const float Ms = (M == -INFINITY) ? 0.0f : M;
#pragma unroll
    for (int w = 0; w < kCoopWarps; ++w) {
        if (LOSS & HP_LOSS_KL) sum_exp += slot->s[w] * ((wm[w] == -INFINITY) ? 0.0f : ex2_approx((wm[w] - Ms) * kLog2e));
        sum_p += slot->sp[w];
        sum_pp += slot->spp[w];
    }
    if (sum_p != sum_p) {
        // a NaN (or +inf with -inf) is in the map: redo the argmax with numpy's exact rules from memory
        ArgMax sx = am_init();
        const float4* m4 = reinterpret_cast<const float4*>(pm);
        for (int e4 = lane; e4 < a.HW / 4; e4 += 32) am_scan4<true>(sx, ldg_stream4(m4 + e4), e4 * 4);
        am = warp_argmax(sx, lane);
        sum_exp = __int_as_float(0x7fc00000);  // log_softmax of a map holding a NaN is NaN
    }
    Centre cc;
    cc.x = slot->cx;
    cc.y = slot->cy;
    const float wgt = slot->weight;
    const PatchSums pst = slot->patch;
    uint32_t qy, qx;
    a.wdiv.divmod(static_cast<uint32_t>(am.i), qy, qx);
    const float keep = (am.v > 0.0f) ? 1.0f : 0.0f;  // NaN -> 0 (keypoint_detection.py:31-34)
    const float px = static_cast<float>(qx) * keep, py = static_cast<float>(qy) * keep;
    const bool pasted = cc.y != kNoPaste;
    // decoding the generated target: its unique maximum (exactly 1.0) sits on the centre when pasted;
    // the all-zero map decodes to the masked (0,0)   
    const float tx = pasted ? static_cast<float>(cc.x) : 0.0f, ty = pasted ? static_cast<float>(cc.y) : 0.0f;
    int valid, hit;
    pipe_pck(a, px, py, tx, ty, valid, hit);
    double mse, kl;
    pipe_losses<LOSS>(a, cc, wgt, am.v, sum_exp, sum_p, sum_pp, pst, mse, kl);
    if (lane == 0) {
        a.pred_xy[2 * map + 0] = px;
        a.pred_xy[2 * map + 1] = py;
        if (a.maxvals) a.maxvals[map] = am.v;
        if (a.weight_out) a.weight_out[map] = wgt;
        const int k = map % a.K;
        if (valid) atomicAdd(&a.ws->counts[a.K + k], 1);
        if (hit) atomicAdd(&a.ws->counts[k], 1);
        if (LOSS & HP_LOSS_MSE) block_loss_add(bl, 0, mse);
        if (LOSS & HP_LOSS_KL) block_loss_add(bl, 1, kl);
    }
}

template <int NV, int LOSS>
__device__ __forceinline__ void pipeline_coop_body(const PipeArgs& a) {
    extern __shared__ float s_tab[];
    __shared__ CoopSlot s_slot[2];
    __shared__ BlockLoss s_loss;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int stride = static_cast<int>(gridDim.x);
    int map = static_cast<int>(blockIdx.x);

    float4 bufA[NV], bufB[NV];
    double jxA = 0.0, jyA = 0.0, jxB = 0.0, jyB = 0.0;
    float visA = 0.f, visB = 0.f;
    if (map < a.n_maps) {
        coop_load<NV>(a.pred, map, a.HW, warp, lane, bufA);
        jxA = a.joints[2 * map];
        jyA = a.joints[2 * map + 1];
        visA = a.vis[map];
    }
    load_table(s_tab, a.tab, a.tmp);
    if (threadIdx.x == 0) block_loss_zero(&s_loss);
    __syncthreads();

    // Two maps per trip so the two register buffers have static names.  The finisher of a map updates
    // s_loss after that map's barrier and before it reaches the next map's barrier, so updates are ordered.
    for (int iter = 0; map < a.n_maps; iter += 2) {
        int next = map + stride;
        if (next < a.n_maps) {
            coop_load<NV>(a.pred, next, a.HW, warp, lane, bufB);
            jxB = a.joints[2 * next];
            jyB = a.joints[2 * next + 1];
            visB = a.vis[next];
        }
        coop_map<NV, LOSS>(a, bufA, map, iter, warp, lane, jxA, jyA, visA, s_tab, &s_slot[0], &s_loss);
        map = next;
        if (map >= a.n_maps) break;
        next = map + stride;
        if (next < a.n_maps) {
            coop_load<NV>(a.pred, next, a.HW, warp, lane, bufA);
            jxA = a.joints[2 * next];
            jyA = a.joints[2 * next + 1];
            visA = a.vis[next];
        }
        coop_map<NV, LOSS>(a, bufB, map, iter + 1, warp, lane, jxB, jyB, visB, s_tab, &s_slot[1], &s_loss);
        map = next;
    }
    __syncthreads();
    if (threadIdx.x == 0) block_loss_flush(&s_loss, a.ws);
    if (pipeline_last_block(a.ws)) {
        if (threadIdx.x == 0) pipeline_publish(a);
    }
}

// two residencies of the same body: 4 blocks/SM (<=128 registers) and 3 blocks/SM (<=168 registers)
template <int NV, int LOSS>
__global__ void __launch_bounds__(32 * kCoopWarps, 4) pipeline_coop4_kernel(const PipeArgs a) {
    pipeline_coop_body<NV, LOSS>(a);
}
template <int NV, int LOSS>
__global__ void __launch_bounds__(32 * kCoopWarps, 3) pipeline_coop3_kernel(const PipeArgs a) {
    pipeline_coop_body<NV, LOSS>(a);
}

}  // namespace hp
