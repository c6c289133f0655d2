// hp_regdisp_sparse.cuh - RegressionDisparityx1 / x5 (and x2 / x3 / x4, which share the x1 recipe) forward, both modes
// (x5 'max' also WITH the fused map y_adv2, see the FUSED instantiation below), as ONE kernel: the pseudo-label decode of y[b,k] (64x64 or any 4096-pixel map) and the KL loss of
// the LOW-RESOLUTION head y_adv[b,k] (32x32: x5, PseudoLabelGenerator03; 16x16: x1, PseudoLabelGenerator01) against
//   'min':  gt = Gaussian at (decoded centre >> shift)                          regda_7.py:3250-3268, 3529-3561
//   'max':  gf = clip(1 - 10 gt)  (its maximum is 1: no normalisation needed)   regda_7.py:3036-3037, 3198-3199
// Both targets are a CONSTANT outside the joint's own patch (0 / 1), so the loss needs no cross-joint information and the
// two launches of the general path (decode of y into `centres`: 29 us per 512 x 21 maps, then the staged loss kernel over
// the small maps) collapse into the shape of hp_regdisp_min.cuh - with one difference: the head is 4 KB / 1 KB per map,
// so it gets its OWN buffer behind the 16 KB stage of y and both copies are requested together under one mbarrier
// (re-using the stage, as the 64x64 kernel does, would put the latency of a tiny copy on every map).
// Algorithmic bytes per map: H*W*4 (y) + oh*ow*4 (y_adv) + 4 (weight) + 24 written.  Roofline: HBM.
#pragma once
#include "hp_regdisp_min.cuh"

namespace hp {

struct RDSparseArgs {
    RDMinArgs m;        // y, y_adv, weight, n_maps, B, K, H, W (of y), HW, tmp, wdiv (by W of y), sdiv, tab, eps, outputs
    int oh, ow, shift;  // the head's size, centre = decoded coordinate >> shift
    float bg;           // the target outside the own patch: 0 ('min') or 1 ('max')
    int want_gf;        // 'max': the patch holds clip(1 - 10 t)
    const float* fused; // x5 'max' with the fused map y_adv2 (train1.py:421 `target0`): g = clip(clip(1 - 10 t) + f - 100 t),
                        // divided by its maximum (regda_7.py:3548-3553) - FUSED instantiation only
    float ubg, ubg_log_ubg;  // bg + eps, (bg + eps) ln (bg + eps)
};

// NITA: iterations (of 32 float4) over the head: oh*ow / 128.  FUSED: a third buffer receives the fused map; the target
// then depends on the pixel everywhere (clip(1 + f) outside the patch), so the head takes two passes: g (written over f)
// and its maximum, then the sums of u = g / M + eps - the recipe of hp_regdisp_dense.cuh with a per-sample label of 1.
template <int W, int BPS, int NITA, bool FUSED>
__global__ void __launch_bounds__(32 * W, BPS) regdisp_sparse_kernel(const RDSparseArgs s) {
    extern __shared__ __align__(128) unsigned char s_rds2[];
    __shared__ PatchSlot s_patch[kTileMaxPatch * 32];
    __shared__ unsigned long long s_acc[kFxAccWords];
    const RDMinArgs& a = s.m;
    constexpr int NITC = 32;
    constexpr uint32_t kYBytes = NITC * 512, kABytes = NITA * 512, kStage = kYBytes + (FUSED ? 2 : 1) * kABytes;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_local = (a.n_maps - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
    const int n_mine = (n_local > warp) ? (n_local - warp + W - 1) / W : 0;
    unsigned char* my_stage = s_rds2 + static_cast<size_t>(warp) * kStage;
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_rds2 + static_cast<size_t>(W) * kStage) + warp;
    const uint32_t stage_u32 = smem_addr(my_stage), bar_u32 = smem_addr(bars);
    const size_t first_map = static_cast<size_t>(blockIdx.x) + static_cast<size_t>(warp) * gridDim.x;
    const size_t map_step = static_cast<size_t>(gridDim.x) * W;  // between consecutive maps of this warp
    const int ohw = s.oh * s.ow;
    const uint64_t pol = l2_evict_first_policy();
    auto request = [&](size_t map) {  // lane 0: y[map] and y_adv[map] under one barrier
        mbar_arrive_expect_tx(bar_u32, kStage);
        bulk_load(stage_u32, a.y + map * a.HW, kYBytes, bar_u32, pol);
        bulk_load(stage_u32 + kYBytes, a.y_adv + map * ohw, kABytes, bar_u32, pol);
        if (FUSED) bulk_load(stage_u32 + kYBytes + kABytes, s.fused + map * ohw, kABytes, bar_u32, pol);
    };

    // ---- prologue: barrier, the small loads, the first copies, then the patch table (order: see hp_pipeline_bulk.cuh) ----
    if (lane == 0) {
        mbar_init(bar_u32, 1);
        mbar_init_fence();
    }
    float w_cur = 1.0f, w_nxt = 1.0f;  // lane l: the weight of the warp's map 32 * batch + l, fetched a batch ahead
    auto load_weights = [&](int batch, float& w) {
        const int jj = batch * 32 + lane;
        if (a.weight && jj < n_mine) w = a.weight[static_cast<int>(blockIdx.x) + (warp + jj * W) * static_cast<int>(gridDim.x)];
    };
    load_weights(0, w_cur);
    if (lane == 0 && n_mine > 0) request(first_map);
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
            const float t = a.tab[sl.dx * sl.dx + sl.dy * sl.dy];
            // the patch's target value: the Gaussian ('min') or clip(1 - 10 t) ('max'; regda_7.py:3036, :3198)
            sl.t = (s.want_gf && !FUSED) ? clip01(__fsub_rn(1.0f, __fmul_rn(t, 10.0f))) : t;
            const float u = sl.t + a.eps;
            sl.ulogu = (u != 0.0f) ? u * logf(u) : 0.0f;
        }
        s_patch[i] = sl;
    }
    if (threadIdx.x < kFxAccWords) s_acc[threadIdx.x] = 0ull;
    __syncthreads();

    uint32_t parity = 0;
    for (int jj = 0; jj < n_mine; ++jj) {
        const size_t map_l = first_map + static_cast<size_t>(jj) * map_step;
        const int map = static_cast<int>(map_l);
        const float4* ybuf = reinterpret_cast<const float4*>(my_stage);
        const float4* abuf = reinterpret_cast<const float4*>(my_stage + kYBytes);
        if (jj != 0 && (jj & 31) == 0) {
            w_cur = w_nxt;
            load_weights((jj >> 5) + 1, w_nxt);
        }
        const float weight = __shfl_sync(0xffffffffu, w_cur, jj & 31);
        mbar_wait(bar_u32, parity);
        parity ^= 1u;
        // ---- y[b,k]: maximum and its first index (numpy argmax rules), nothing else -------------------------------------
        float run = -INFINITY;
        int best_it = 0;
        float2 chk = make_float2(0.f, 0.f);  // a NaN anywhere in the map poisons this sum
#pragma unroll 8
        for (int it = 0; it < NITC; ++it) {
            const float4 v = ybuf[it * 32 + lane];
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
            const float4 v = ybuf[kmin & 1023u];
            const int comp = (v.x == am.v) ? 0 : ((v.y == am.v) ? 1 : ((v.z == am.v) ? 2 : 3));
            am.i = static_cast<int>(kmin & 1023u) * 4 + comp;
        }
        const float chk_all = warp_sum(chk.x + chk.y);
        if (chk_all != chk_all) {
            // a NaN (or +inf with -inf) is in y[b,k]: redo the argmax with numpy's exact rules from memory (L2)
            ArgMax sx = am_init();
            const float4* m4 = reinterpret_cast<const float4*>(a.y + map_l * a.HW);
            for (int e4 = lane; e4 < a.HW / 4; e4 += 32) am_scan4<true>(sx, ldg_stream4(m4 + e4), e4 * 4);
            am = warp_argmax(sx, lane);
        }
        // ---- the pseudo label: patch at the decoded (masked) centre >> shift (regda_7.py:3033, :3195) ---------------------
        uint32_t qy, qx;
        a.wdiv.divmod(static_cast<uint32_t>(am.i), qy, qx);
        const bool keep = am.v > 0.0f;  // NaN -> (0, 0)  (keypoint_detection.py:31-34)
        const int cx = keep ? static_cast<int>(qx) >> s.shift : 0, cy = keep ? static_cast<int>(qy) >> s.shift : 0;
        if (FUSED) {
            float* F = reinterpret_cast<float*>(my_stage + kYBytes + kABytes);
            float4* F4 = reinterpret_cast<float4*>(F);
            const float* P = reinterpret_cast<const float*>(abuf);
            // ---- own patch: exact un-normalised values into registers, -inf into the fused map (the passes see g = 0) ----
            float gex[kTileMaxPatch], pk[kTileMaxPatch];
            int poff[kTileMaxPatch];
            float mg = -INFINITY;
#pragma unroll
            for (int k = 0; k < kTileMaxPatch; ++k) {
                const PatchSlot sl = s_patch[k * 32 + lane];
                const int x = cx + sl.dx, yy = cy + sl.dy;
                const bool in = static_cast<unsigned>(x) < static_cast<unsigned>(s.ow) && static_cast<unsigned>(yy) < static_cast<unsigned>(s.oh);
                const int off = in ? yy * s.ow + x : 0;
                float g = clip01(__fsub_rn(1.0f, __fmul_rn(sl.t, 10.0f)));
                g = clip01(__fsub_rn(__fadd_rn(g, F[off]), __fmul_rn(sl.t, 100.0f)));
                gex[k] = g;
                pk[k] = P[off];
                poff[k] = in ? off : -1;
                mg = in ? fmaxf(mg, g) : mg;
            }
            __syncwarp();
#pragma unroll
            for (int k = 0; k < kTileMaxPatch; ++k)
                if (poff[k] >= 0) F[poff[k]] = -INFINITY;
            __syncwarp();
            // ---- pass 1: g = clip(1 + f) written over f, its maximum, the softmax maximum ------------------------------------
            float run = -INFINITY;
#pragma unroll
            for (int it = 0; it < NITA; ++it) {
                const float4 f = F4[it * 32 + lane], v = abuf[it * 32 + lane];
                const float4 g = clip01_4(add4(make_float4(1.f, 1.f, 1.f, 1.f), f));
                F4[it * 32 + lane] = g;
                mg = fmaxf(mg, max4(g));
                run = fmaxf(run, max4(v));
            }
            const float Mp = warp_max_f32(run), M = warp_max_f32(mg);
            const float invM = (M == 1.0f) ? 1.0f : __frcp_rn(M);
            const float ms = (Mp == -INFINITY) ? 0.0f : Mp, mb = -ms * kLog2e;
            // ---- pass 2: the sums of u = g / M + eps, against the softmax maximum in log2 units -----------------------------
            float sexp = 0.f, su = 0.f, sua = 0.f, sulg = 0.f;
#pragma unroll
            for (int it = 0; it < NITA; ++it) {
                const float4 v = abuf[it * 32 + lane], g = F4[it * 32 + lane];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const float av = fmaf(f4_get(v, c), kLog2e, mb);
                    const float u = fmaf(f4_get(g, c), invM, a.eps);
                    sexp += ex2_approx(av);
                    su += u;
                    sua = fmaf(u, av, sua);
                    sulg = fmaf(u, lg2_approx(fmaxf(u, 1.17549435e-38f)), sulg);
                }
            }
            const float u_bg = fmaf(0.0f, invM, a.eps);
            const float bg_ulg = u_bg * lg2_approx(fmaxf(u_bg, 1.17549435e-38f));
#pragma unroll
            for (int k = 0; k < kTileMaxPatch; ++k) {
                const float uex = fmaf(gex[k], invM, a.eps);
                const float du = uex - u_bg;
                if (poff[k] >= 0) {
                    su += du;
                    sua = fmaf(du, fmaf(pk[k], kLog2e, mb), sua);
                    sulg += uex * lg2_approx(fmaxf(uex, 1.17549435e-38f)) - bg_ulg;
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // this lane's writes to the buffer precede the next copy
            __syncwarp();
            if (lane == 0 && jj + 1 < n_mine) request(map_l + map_step);
            const float r = warp_sum3_scattered(sexp, su, sua, lane);
            const float Sexp = __shfl_sync(0xffffffffu, r, 0), Su = __shfl_sync(0xffffffffu, r, 8), Sua = __shfl_sync(0xffffffffu, r, 16);
            const float Sulg = warp_sum(sulg);
            if (lane == 0) {
                const float lg_se = lg2_approx(Sexp);
                const float lse = fmaf(lg_se, kLn2, ms);
                const float L = kLn2 * (__fdividef(Sulg - Sua, Su) - lg2_approx(Su) + lg_se);
                const float Lw = L * weight;
                a.per_map[map] = Lw;
                a.stats[3 * map + 0] = lse;
                a.stats[3 * map + 1] = Su;
                a.stats[3 * map + 2] = M;
                *reinterpret_cast<int2*>(a.centres + 2 * static_cast<size_t>(map)) = make_int2(cx, cy);
                if (a.mean) fx_acc_add(s_acc, Lw);
            }
            continue;
        }
        PatchSums ps{0.f, 0.f, 0.f, 0.f, 0.f};
        const float* fbuf = reinterpret_cast<const float*>(abuf);
#pragma unroll
        for (int k = 0; k < kTileMaxPatch; ++k) {
            const PatchSlot sl = s_patch[k * 32 + lane];
            const int x = cx + sl.dx, yy = cy + sl.dy;
            const bool in = static_cast<unsigned>(x) < static_cast<unsigned>(s.ow) && static_cast<unsigned>(yy) < static_cast<unsigned>(s.oh);
            const float pk = fbuf[in ? yy * s.ow + x : 0];
            const float u = sl.t + a.eps;
            ps.ulogu += in ? sl.ulogu : 0.0f;
            ps.u += in ? u : 0.0f;
            ps.up = in ? fmaf(u, pk, ps.up) : ps.up;
            ps.p += in ? pk : 0.0f;
        }
        // ---- y_adv[b,k]: exact two-pass softmax sums ------------------------------------------------------------------------
        BulkAcc A;
        A.M = -INFINITY;
        A.idx = 0;
        A.s2 = A.sp2 = A.spp2 = make_float2(0.f, 0.f);
        bulk_chunk<NITA, HP_LOSS_KL, false>(A, abuf, 0, lane);
        __syncwarp();
        if (lane == 0 && jj + 1 < n_mine) request(map_l + map_step);  // both buffers have been read out
        float sum_exp, sum_p;
        {
            const float r = warp_sum3_scattered(A.s2.x + A.s2.y, A.sp2.x + A.sp2.y, ps.u, lane);
            sum_exp = __shfl_sync(0xffffffffu, r, 0);
            sum_p = __shfl_sync(0xffffffffu, r, 8);
            ps.u = __shfl_sync(0xffffffffu, r, 16);
        }
        {
            const float r = warp_sum3_scattered(ps.up, ps.p, ps.ulogu, lane);
            ps.up = __shfl_sync(0xffffffffu, r, 0);
            ps.p = __shfl_sync(0xffffffffu, r, 8);
            ps.ulogu = __shfl_sync(0xffffffffu, r, 16);
        }
        // ---- closure: L = (sum u ln u - sum u p)/S - ln S + lse  (loss.py:145-158); outside the patch u = bg + eps ------------
        if (lane == 0) {
            const int area = (min(cx + a.tmp, s.ow - 1) - max(cx - a.tmp, 0) + 1) * (min(cy + a.tmp, s.oh - 1) - max(cy - a.tmp, 0) + 1);
            const float n_bg = static_cast<float>(ohw - area);
            const float Su = fmaf(s.ubg, n_bg, ps.u);
            const float Sup = fmaf(s.ubg, sum_p - ps.p, ps.up);
            const float Sulogu = fmaf(n_bg, s.ubg_log_ubg, ps.ulogu);
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

template <int W, int BPS, int NITA, bool FUSED>
static int launch_rds2_shape(const RDSparseArgs& s, int sms, cudaStream_t stream, const char* who) {
    constexpr size_t smem = static_cast<size_t>(W) * (32 * 512 + (FUSED ? 2 : 1) * NITA * 512) + sizeof(uint64_t) * W;
    static bool configured[16] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 16 || !configured[dev]) {
        const cudaError_t e = cudaFuncSetAttribute(regdisp_sparse_kernel<W, BPS, NITA, FUSED>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                   static_cast<int>(smem));
        if (e != cudaSuccess) return fail(static_cast<int>(e), "%s: %s", who, cudaGetErrorString(e));
        if (dev >= 0 && dev < 16) configured[dev] = true;
    }
    int grid = sms * BPS;
    if (grid > s.m.n_maps) grid = s.m.n_maps;
    regdisp_sparse_kernel<W, BPS, NITA, FUSED><<<grid, 32 * W, smem, stream>>>(s);
    return launch_status(who);
}

// returns 1 when the shape is not covered (the caller takes the general two-launch path), 0 when launched
static int launch_regdisp_sparse(const RDSparseArgs& s, cudaStream_t stream, const char* who) {
    static int sms = 0;
    if (sms == 0) {
        sms = hp_device_sm_count();
        if (sms <= 0) sms = 148;
    }
    const int ohw = s.oh * s.ow;
    if (s.fused) return ohw == 1024 ? launch_rds2_shape<4, 2, 8, true>(s, sms, stream, who) : 1;  // 2 x 4 x 24 KB per SM
    if (ohw == 1024) return launch_rds2_shape<5, 2, 8, false>(s, sms, stream, who);   // 2 x 5 x 20 KB per SM
    if (ohw == 256) return launch_rds2_shape<4, 3, 2, false>(s, sms, stream, who);    // 3 x 4 x 17 KB per SM
    return 1;
}

}  // namespace hp
