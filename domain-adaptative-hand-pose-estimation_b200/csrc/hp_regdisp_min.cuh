// hp_regdisp_min.cuh - RegressionDisparity 'min' at full resolution (base / x6 / rd4 with 64x64 heads; any H x W of
// 4096 pixels) as ONE kernel: the pseudo-label decode of y and the KL loss of y_adv against the Gaussian at the decoded
// centre (regda_4.py:129-143, regda_7.py:3609-3632 with mode='min': `criterion(y_adv, gt, weight)`,
// gt[b,k] = Gaussian at get_max_preds(y)[b,k]).
//
// 'min' needs NO cross-joint information - the target of map (b,k) is a function of y[b,k]'s own argmax - so the two
// launches of the general path (decode of y into `centres`, then the staged loss over y_adv) collapse into the shape of
// the fused pipeline kernel (hp_pipeline_bulk.cuh): one warp per map pair, a private 16 KB stage filled by the copy
// engine; the warp first receives y[b,k] (pass A only: maximum + first index), re-arms the same stage with y_adv[b,k]
// while it derives the centre and the patch offsets, then runs the exact two-pass softmax + patch terms on it.
// Algorithmic bytes per map: 2 * H*W*4 (y, y_adv) + 4 (weight) + 24 written (loss, lse, S, M, centre) = the floor of
// configs[2] 'min' (32,768 B).  The general path costs 29 us (decode) + 44 us (loss, issue-bound at
// 2,480 warp-instructions per map) on 512 x 21 maps; this kernel spends ~1,400 per map PAIR.
#pragma once
#include "hp_pipeline_bulk.cuh"

namespace hp {

struct RDMinArgs {
    const float* y;
    const float* y_adv;
    const float* weight;  // nullable: 1
    int n_maps, B, K, H, W, HW, tmp;
    FastDiv wdiv, sdiv;   // by W ; by the patch side
    const float* tab;
    float eps, eps_log_eps;
    float* per_map;
    float* per_sample;    // nullable
    float* mean;          // nullable
    float* stats;         // [n_maps, 3] lse, S, M (= 1)
    int32_t* centres;     // [n_maps, 2]
    Workspace* ws;
};

template <int W, int BPS>
__global__ void __launch_bounds__(32 * W, BPS) regdisp_min_kernel(const RDMinArgs a) {
    extern __shared__ __align__(128) unsigned char s_rdm[];
    __shared__ PatchSlot s_patch[kTileMaxPatch * 32];
    __shared__ unsigned long long s_acc[kFxAccWords];
    constexpr int NITC = 32;
    constexpr int kBytes = NITC * 512;  // 16 KB: one 4096-pixel map
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_local = (a.n_maps - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
    const int n_mine = (n_local > warp) ? (n_local - warp + W - 1) / W : 0;
    unsigned char* my_stage = s_rdm + static_cast<size_t>(warp) * kBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_rdm + static_cast<size_t>(W) * kBytes) + warp;
    const uint32_t stage_u32 = smem_addr(my_stage), bar_u32 = smem_addr(bars);
    const size_t map_stride = static_cast<size_t>(gridDim.x) * W * a.HW;  // between consecutive maps of this warp
    const size_t first_off = (static_cast<size_t>(blockIdx.x) + static_cast<size_t>(warp) * gridDim.x) * a.HW;

    // ---- prologue: barrier, the small loads, the first copy, then the patch table (order: see hp_pipeline_bulk.cuh) ----
    if (lane == 0) {
        mbar_init(bar_u32, 1);
        mbar_init_fence();
    }
    // weights are lane-distributed: lane l holds the weight of the warp's map 32*batch + l, fetched a batch ahead
    float w_cur = 1.0f, w_nxt = 1.0f;
    auto load_weights = [&](int batch, float& w) {
        const int jj = batch * 32 + lane;
        if (a.weight && jj < n_mine) w = a.weight[static_cast<int>(blockIdx.x) + (warp + jj * W) * static_cast<int>(gridDim.x)];
    };
    load_weights(0, w_cur);
    const uint64_t pol = l2_evict_first_policy();
    if (lane == 0 && n_mine > 0) {
        mbar_arrive_expect_tx(bar_u32, kBytes);
        bulk_load(stage_u32, a.y + first_off, kBytes, bar_u32, pol);
    }
    load_weights(1, w_nxt);
    const int side = 2 * a.tmp + 1, n_patch = side * side;
    for (int i = threadIdx.x; i < kTileMaxPatch * 32; i += 32 * W) {
        PatchSlot sl;
        sl.dx = 1 << 20; sl.dy = 0; sl.t = 0.f; sl.ulogu = 0.f;
        if (i < n_patch) {
            uint32_t ry, rx;
            a.sdiv.divmod(static_cast<uint32_t>(i), ry, rx);
            sl.dx = static_cast<int>(rx) - a.tmp;
            sl.dy = static_cast<int>(ry) - a.tmp;
            sl.t = a.tab[sl.dx * sl.dx + sl.dy * sl.dy];
            const float u = sl.t + a.eps;
            sl.ulogu = (u != 0.0f) ? u * logf(u) : 0.0f;
        }
        s_patch[i] = sl;
    }
    if (threadIdx.x < kFxAccWords) s_acc[threadIdx.x] = 0ull;
    __syncthreads();

    uint32_t parity = 0;
    for (int jj = 0; jj < n_mine; ++jj) {
        const int map = static_cast<int>(blockIdx.x) + (warp + jj * W) * static_cast<int>(gridDim.x);
        const size_t off = first_off + static_cast<size_t>(jj) * map_stride;
        const float4* buf = reinterpret_cast<const float4*>(my_stage);
        if (jj != 0 && (jj & 31) == 0) {
            w_cur = w_nxt;
            load_weights((jj >> 5) + 1, w_nxt);
        }
        const float weight = __shfl_sync(0xffffffffu, w_cur, jj & 31);
        // ---- y[b,k]: maximum and its first index (numpy argmax rules), nothing else ---------------------------------
        mbar_wait(bar_u32, parity);
        parity ^= 1u;
        float run = -INFINITY;
        int best_it = 0;
        float2 chk = make_float2(0.f, 0.f);  // a NaN anywhere in the map poisons this sum
#pragma unroll 8
        for (int it = 0; it < NITC; ++it) {
            const float4 v = buf[it * 32 + lane];
            const float t = fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w));
            best_it = (t > run) ? it : best_it;  // strict: the earlier iteration keeps ties
            run = fmaxf(run, t);
            chk = __fadd2_rn(chk, __fadd2_rn(make_float2(v.x, v.y), make_float2(v.z, v.w)));
        }
        ArgMax am;
        am.v = warp_max_f32(run);
        {
            const unsigned key = (run == am.v) ? static_cast<unsigned>(best_it * 32 + lane) : 0x7fffffffu;
            const unsigned kmin = __reduce_min_sync(0xffffffffu, key);
            const float4 v = buf[kmin & 1023u];
            const int comp = (v.x == am.v) ? 0 : ((v.y == am.v) ? 1 : ((v.z == am.v) ? 2 : 3));
            am.i = static_cast<int>(kmin & 1023u) * 4 + comp;
        }
        const float chk_all = warp_sum(chk.x + chk.y);
        // the stage has been read out: request y_adv[b,k] into it
        __syncwarp();
        if (lane == 0) {
            mbar_arrive_expect_tx(bar_u32, kBytes);
            bulk_load(stage_u32, a.y_adv + off, kBytes, bar_u32, pol);
        }
        if (chk_all != chk_all) {
            // a NaN (or +inf with -inf) is in y[b,k]: redo the argmax with numpy's exact rules from memory (L2)
            ArgMax sx = am_init();
            const float4* m4 = reinterpret_cast<const float4*>(a.y + off);
            for (int e4 = lane; e4 < a.HW / 4; e4 += 32) am_scan4<true>(sx, ldg_stream4(m4 + e4), e4 * 4);
            am = warp_argmax(sx, lane);
        }
        // ---- the pseudo label: Gaussian at the decoded (masked) centre; everything that does not need y_adv ----------
        uint32_t qy, qx;
        a.wdiv.divmod(static_cast<uint32_t>(am.i), qy, qx);
        const bool keep = am.v > 0.0f;  // NaN -> (0, 0)  (keypoint_detection.py:31-34)
        const int cx = keep ? static_cast<int>(qx) : 0, cy = keep ? static_cast<int>(qy) : 0;
        PatchSums ps{0.f, 0.f, 0.f, 0.f, 0.f};
        int poff[kTileMaxPatch];
        float tv[kTileMaxPatch];
#pragma unroll
        for (int k = 0; k < kTileMaxPatch; ++k) {
            const PatchSlot sl = s_patch[k * 32 + lane];
            const int x = cx + sl.dx, yy = cy + sl.dy;
            const bool in = static_cast<unsigned>(x) < static_cast<unsigned>(a.W) && static_cast<unsigned>(yy) < static_cast<unsigned>(a.H);
            poff[k] = in ? yy * a.W + x : -1;
            tv[k] = in ? sl.t : 0.0f;
            ps.ulogu += in ? sl.ulogu : 0.0f;
            ps.u += (tv[k] != 0.0f) ? tv[k] + a.eps : 0.0f;
        }
        {
            const float r = warp_sum3_scattered(ps.u, ps.ulogu, 0.0f, lane);
            ps.u = __shfl_sync(0xffffffffu, r, 0);
            ps.ulogu = __shfl_sync(0xffffffffu, r, 8);
        }
        // ---- y_adv[b,k]: exact two-pass softmax sums + the patch terms ---------------------------------------------------
        BulkAcc A;
        A.M = -INFINITY;
        A.idx = 0;
        A.s2 = A.sp2 = A.spp2 = make_float2(0.f, 0.f);
        mbar_wait(bar_u32, parity);
        parity ^= 1u;
        bulk_chunk<NITC, HP_LOSS_KL, false>(A, buf, 0, lane);
        const float* fbuf = reinterpret_cast<const float*>(my_stage);
#pragma unroll
        for (int k = 0; k < kTileMaxPatch; ++k) {
            const bool here = poff[k] >= 0;
            const float pk = here ? fbuf[poff[k]] : 0.0f;
            const float tk = here ? tv[k] : 0.0f;
            const float u = (tk != 0.0f) ? tk + a.eps : 0.0f;
            ps.up = fmaf(u, pk, ps.up);
            ps.p += pk;
        }
        __syncwarp();
        if (lane == 0 && jj + 1 < n_mine) {  // the next pair's y
            mbar_arrive_expect_tx(bar_u32, kBytes);
            bulk_load(stage_u32, a.y + off + map_stride, kBytes, bar_u32, pol);
        }
        float sum_exp, sum_p;
        {
            const float r = warp_sum3_scattered(A.s2.x + A.s2.y, A.sp2.x + A.sp2.y, 0.0f, lane);
            sum_exp = __shfl_sync(0xffffffffu, r, 0);
            sum_p = __shfl_sync(0xffffffffu, r, 8);
        }
        {
            const float r = warp_sum3_scattered(ps.up, ps.p, 0.0f, lane);
            ps.up = __shfl_sync(0xffffffffu, r, 0);
            ps.p = __shfl_sync(0xffffffffu, r, 8);
        }
        // ---- closure: L = (sum u ln u - sum u p)/S - ln S + lse  (loss.py:145-158), u = t + eps, background u = eps --------
        if (lane == 0) {
            const int area = (min(cx + a.tmp, a.W - 1) - max(cx - a.tmp, 0) + 1) * (min(cy + a.tmp, a.H - 1) - max(cy - a.tmp, 0) + 1);
            const float n_bg = static_cast<float>(a.HW - area);
            const float Su = fmaf(a.eps, n_bg, ps.u);
            const float Sup = fmaf(a.eps, sum_p - ps.p, ps.up);
            const float Sulogu = fmaf(n_bg, a.eps_log_eps, ps.ulogu);
            const float lse = fmaf(lg2_approx(sum_exp), kLn2, A.M);
            const float L = __fdividef(Sulogu - Sup, Su) - lg2_approx(Su) * kLn2 + lse;
            const float Lw = L * weight;
            a.per_map[map] = Lw;
            a.stats[3 * map + 0] = lse;
            a.stats[3 * map + 1] = Su;
            a.stats[3 * map + 2] = 1.0f;
            *reinterpret_cast<int2*>(a.centres + 2 * static_cast<size_t>(map)) = make_int2(cx, cy);
            if (a.mean) fx_acc_add(s_acc, Lw);
        }
    }

    // ---- epilogue: block sum -> workspace, the last block finalises 'mean' / the per-sample means ----------------------
    if (a.mean == nullptr && a.per_sample == nullptr) return;
    __syncthreads();
    if (a.mean && threadIdx.x < kFxAccWords && s_acc[threadIdx.x] != 0ull) atomicAdd(&a.ws->acc[threadIdx.x], s_acc[threadIdx.x]);
    if (last_block_arrives(&a.ws->counter, gridDim.x)) {
        if (a.per_sample) per_sample_means(a.per_map, a.B, a.K, a.per_sample, threadIdx.x, 32 * W);
        if (a.mean && threadIdx.x == 0) *a.mean = fx_mean_from_workspace(a.ws->acc, a.n_maps);
        if (threadIdx.x == 0) a.ws->counter = 0;
    }
}

// returns 1 when the shape is not covered (the caller takes the general two-launch path), 0 when launched
static int launch_regdisp_min(const RDMinArgs& a, cudaStream_t stream, const char* who) {
    constexpr int W = 4, BPS = 3;
    constexpr size_t smem = static_cast<size_t>(W) * 32 * 512 + sizeof(uint64_t) * W;
    static int sms = 0;
    if (sms == 0) {
        sms = hp_device_sm_count();
        if (sms <= 0) sms = 148;
    }
    static bool configured[16] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 16 || !configured[dev]) {
        const cudaError_t e = cudaFuncSetAttribute(regdisp_min_kernel<W, BPS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                   static_cast<int>(smem));
        if (e != cudaSuccess) return fail(static_cast<int>(e), "%s: %s", who, cudaGetErrorString(e));
        if (dev >= 0 && dev < 16) configured[dev] = true;
    }
    int grid = sms * BPS;
    if (grid > a.n_maps) grid = a.n_maps;
    regdisp_min_kernel<W, BPS><<<grid, 32 * W, smem, stream>>>(a);
    return launch_status(who);
}

}  // namespace hp
