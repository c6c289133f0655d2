// hp_pipeline.cu - the benchmarked pipeline: target generation + MSE + KL + decode + PCK in ONE
// pass over the prediction tensor (BASELINE.json metric: heatmaps/s, gen+loss+decode+PCK).
//
// Replaces, per batch:  generate_target x B      uda/dataset/util.py:9-68
//                       JointsMSELoss            uda/model/loss.py:55-65
//                       JointsKLLoss             uda/model/loss.py:145-158
//                       accuracy                 utils/keypoint_detection.py:63-92 (2x get_max_preds + PCK)
// The Gaussian target of a map is a function of its centre only, so it is regenerated in
// registers (table lookup on dx^2+dy^2) next to the prediction values and never touches memory;
// decoding the generated target is exactly (mu_x, mu_y) when pasted and (0,0) otherwise
// (SURVEY.md appendix A4), so PCK needs no second argmax.
//
// HBM layout: pred [B,K,H,W] fp32 contiguous; joints fp64 [B*K,2]; vis fp32 [B*K].
// Algorithmic bytes per map: H*W*4 (pred) + 24 (joint, vis, weight) + 8 (coords) = 16,416 at 64x64.
// Roofline: HBM.  Per element: 1 compare-select pair (argmax), 1 FFMA+MUFU+FADD (softmax),
// 1 FADD (sum p), 1 FFMA (squared error); target terms only on the 13 rows the patch touches.
#include <cmath>

#include "hp_common.cuh"
#include "hp_dispatch.cuh"

namespace hp {

struct PipeArgs {
    const float* pred;
    const double* joints;
    const float* vis;
    int n_maps, K, H, W;
    FastDiv wdiv;
    double sx, sy;
    int tmp;
    const float* tab;
    float eps;
    double thr;
    int loss_mask;
    float* pred_xy;
    float* maxvals;
    float* weight_out;
    double* partial;
    int accumulate;
    double* result;
    Workspace* ws;
    float* map_vals;  // [2*n_maps] per-map mse / kl (workspace tail)
};

// number of in-bounds pixels of the pasted patch
__device__ __forceinline__ int patch_area(Centre c, int tmp, int W, int H) {
    if (c.y == kNoPaste) return 0;
    const int nx = min(c.x + tmp, W - 1) - max(c.x - tmp, 0) + 1;
    const int ny = min(c.y + tmp, H - 1) - max(c.y - tmp, 0) + 1;
    return nx * ny;
}

__device__ __forceinline__ void pipeline_result_from_partial(const double* p, int K, double* result) {
    // p = { mse_sum, kl_sum, n_maps, n_elems, hits[K], valid[K] } ; result = { mse, kl, avg_acc, cnt, acc[K] }
    result[0] = p[0] / p[2];
    result[1] = p[1] / p[2];
    int hits[HP_MAX_K], valid[HP_MAX_K];
    for (int k = 0; k < K; ++k) {
        hits[k] = static_cast<int>(p[4 + k]);
        valid[k] = static_cast<int>(p[4 + K + k]);
    }
    double acc[HP_MAX_K + 2];
    pck_finalize_serial(hits, valid, K, acc);
    for (int k = 0; k < K; ++k) result[4 + k] = acc[k];
    result[2] = acc[K];
    result[3] = acc[K + 1];
}

}  // namespace hp

#include "hp_pipeline_stream.cuh"

namespace hp {

template <int TPM, int NV, int MODE, int MPB>
__global__ void __launch_bounds__(TPM* MPB) pipeline_kernel(const PipeArgs a) {
    // per-map sums: 0 sum (p-t)^2 | 1 sum p | 2 sum_patch u*p | 3 sum_patch u*log(u) | 4 sum_patch u | 5 sum_patch p
    // with u = t + eps; "patch" = the pixels the pasted Gaussian covers (t != 0), everything else is
    // background where u == eps exactly, so its contribution is added in closed form at the end.
    constexpr int NS = 6;
    extern __shared__ float s_tab[];
    __shared__ Stats<NS> scratch[TPM > 32 ? TPM / 32 + 1 : 1];
    __shared__ Centre s_centre[MPB];
    __shared__ float s_weight[MPB];
    __shared__ double s_red[TPM * MPB];

    const int HW = a.H * a.W;
    const int g = threadIdx.x / TPM, t = threadIdx.x % TPM;
    const int map = blockIdx.x * MPB + g;
    const bool active = map < a.n_maps;
    const bool want_mse = (a.loss_mask & HP_LOSS_MSE) != 0, want_kl = (a.loss_mask & HP_LOSS_KL) != 0;
    const float* pm = a.pred + static_cast<size_t>(active ? map : 0) * HW;
    const int ntiles = (MODE == WALK_EXACT) ? 1 : tiles_for<TPM, NV>(HW);

    // issue the first tile's loads before anything else so they are in flight during the setup
    float4 p[NV];
    if (active) load_tile<TPM, NV, MODE>(pm, HW, 0, t, -INFINITY, p);

    load_table(s_tab, a.tab, a.tmp);
    if (active && t == 0) {
        float w;
        s_centre[g] = target_centre(a.joints[2 * map], a.joints[2 * map + 1], a.vis[map], a.sx, a.sy, a.W, a.H, w);
        s_weight[g] = w;
    }
    __syncthreads();

    if (active) {
        const Centre c = s_centre[g];
        Stats<NS> st;
        stats_init(st);
        for (int tile = 0; tile < ntiles; ++tile) {
            if (tile > 0) load_tile<TPM, NV, MODE>(pm, HW, tile, t, -INFINITY, p);
            if (want_kl) softmax_tile<NV>(st.m, st.s, p);
#pragma unroll
            for (int j = 0; j < NV; ++j) {
                const int idx0 = tile * (TPM * NV * 4) + (j * TPM + t) * 4;
                if (MODE != WALK_EXACT && idx0 >= HW) continue;
                uint32_t y0, x0;
                a.wdiv.divmod(static_cast<uint32_t>(idx0), y0, x0);
                float tv[4] = {0.f, 0.f, 0.f, 0.f};
                bool row_hit;
                if (MODE == WALK_SCALAR) {
                    row_hit = true;  // four consecutive elements may straddle rows: test each one
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) {
                        int x = static_cast<int>(x0) + cc, y = static_cast<int>(y0);
                        while (x >= a.W) {
                            x -= a.W;
                            ++y;
                        }
                        tv[cc] = patch_at(s_tab, a.tmp, c, x, y);
                    }
                } else {
                    const int dy = static_cast<int>(y0) - c.y;
                    row_hit = static_cast<unsigned>(dy + a.tmp) <= 2u * static_cast<unsigned>(a.tmp);
                    if (row_hit) {
                        const float4 t4 = patch_at4(s_tab, a.tmp, c, static_cast<int>(x0), static_cast<int>(y0));
                        tv[0] = t4.x; tv[1] = t4.y; tv[2] = t4.z; tv[3] = t4.w;
                    }
                }
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) {
                    if (MODE != WALK_EXACT && idx0 + cc >= HW) continue;
                    const float pv = f4_get(p[j], cc);
                    am_scan1<false>(st.am, pv, idx0 + cc);
                    st.sum[1] += pv;
                    if (want_mse) {
                        const float d = pv - tv[cc];
                        st.sum[0] = fmaf(d, d, st.sum[0]);
                    }
                }
                if (row_hit && want_kl) {
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) {
                        if (tv[cc] != 0.0f && (MODE == WALK_EXACT || idx0 + cc < HW)) {
                            const float pv = f4_get(p[j], cc);
                            const float u = tv[cc] + a.eps;
                            st.sum[2] = fmaf(u, pv, st.sum[2]);
                            st.sum[3] = fmaf(u, __logf(u), st.sum[3]);
                            st.sum[4] += u;
                            st.sum[5] += pv;
                        }
                    }
                }
            }
        }
        group_reduce<TPM, NS, true, true, false>(st, scratch);
        if (st.sum[1] != st.sum[1]) {
            // a NaN (or +inf with -inf) is in the map: redo the argmax with the exact numpy rules
            Stats<NS> sx;
            stats_init(sx);
            for (int tile = 0; tile < ntiles; ++tile) {
                if (MODE != WALK_EXACT) load_tile<TPM, NV, MODE>(pm, HW, tile, t, -INFINITY, p);
#pragma unroll
                for (int j = 0; j < NV; ++j) {
                    const int idx0 = tile * (TPM * NV * 4) + (j * TPM + t) * 4;
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc)
                        if (MODE == WALK_EXACT || idx0 + cc < HW) am_scan1<true>(sx.am, f4_get(p[j], cc), idx0 + cc);
                }
            }
            group_reduce<TPM, NS, true, false, false>(sx, scratch);
            st.am = sx.am;
        }
        if (t == 0) {
            const float w = s_weight[g];
            float px, py;
            decode_xy(st.am, a.W, px, py);
            a.pred_xy[2 * map + 0] = px;
            a.pred_xy[2 * map + 1] = py;
            if (a.maxvals) a.maxvals[map] = st.am.v;
            if (a.weight_out) a.weight_out[map] = w;
            // decoding the generated target: its unique maximum (exactly 1.0) sits on the centre when
            // pasted, and the all-zero map decodes to the masked (0,0)   
            const bool pasted = c.y != kNoPaste;
            const float tx = pasted ? static_cast<float>(c.x) : 0.0f, ty = pasted ? static_cast<float>(c.y) : 0.0f;
            int valid, hit;
            pck_one(px, py, tx, ty, a.H, a.W, a.thr, valid, hit);
            const int k = map % a.K;
            if (valid) atomicAdd(&a.ws->counts[a.K + k], 1);
            if (hit) atomicAdd(&a.ws->counts[k], 1);
            float mse = 0.f, kl = 0.f;
            if (want_mse)  // mean over HW of 0.5*w*(p-t)^2   (loss.py:59-65)
                mse = static_cast<float>(0.5 * static_cast<double>(w) * static_cast<double>(st.sum[0]) /
                                         static_cast<double>(HW));
            if (want_kl) {
                const double eps = static_cast<double>(a.eps);
                const double n_bg = static_cast<double>(HW - patch_area(c, a.tmp, a.W, a.H));
                const double Su = static_cast<double>(st.sum[4]) + eps * n_bg;
                const double Sup = static_cast<double>(st.sum[2]) +
                                   eps * (static_cast<double>(st.sum[1]) - static_cast<double>(st.sum[5]));
                const double Sulogu = static_cast<double>(st.sum[3]) + ((a.eps > 0.0f) ? n_bg * eps * log(eps) : 0.0);
                const double lse = static_cast<double>(st.m) + log(static_cast<double>(st.s));
                const double L = (Sulogu - Sup) / Su - log(Su) + lse;  // Su == 0 (eps 0, nothing pasted) -> NaN
                kl = static_cast<float>(L * static_cast<double>(w));
            }
            a.map_vals[map] = mse;
            a.map_vals[a.n_maps + map] = kl;
        }
    }

    if (last_block_arrives(&a.ws->counter, gridDim.x)) {
        const volatile float* mv = a.map_vals;
        double acc_m = 0.0, acc_k = 0.0;
        for (int i = threadIdx.x; i < a.n_maps; i += TPM * MPB) {
            acc_m += static_cast<double>(mv[i]);
            acc_k += static_cast<double>(mv[a.n_maps + i]);
        }
        s_red[threadIdx.x] = acc_m;
        __syncthreads();
        for (int o = (TPM * MPB) / 2; o > 0; o >>= 1) {
            if (threadIdx.x < o) s_red[threadIdx.x] += s_red[threadIdx.x + o];
            __syncthreads();
        }
        const double mse_sum = s_red[0];
        __syncthreads();
        s_red[threadIdx.x] = acc_k;
        __syncthreads();
        for (int o = (TPM * MPB) / 2; o > 0; o >>= 1) {
            if (threadIdx.x < o) s_red[threadIdx.x] += s_red[threadIdx.x + o];
            __syncthreads();
        }
        const double kl_sum = s_red[0];
        if (threadIdx.x == 0) {
            volatile int* cnt = a.ws->counts;
            double* P = a.partial;
            P[0] = (a.accumulate ? P[0] : 0.0) + mse_sum;
            P[1] = (a.accumulate ? P[1] : 0.0) + kl_sum;
            P[2] = (a.accumulate ? P[2] : 0.0) + static_cast<double>(a.n_maps);
            P[3] = (a.accumulate ? P[3] : 0.0) + static_cast<double>(a.n_maps) * static_cast<double>(HW);
            for (int k = 0; k < 2 * a.K; ++k) {
                P[4 + k] = (a.accumulate ? P[4 + k] : 0.0) + static_cast<double>(cnt[k]);
                cnt[k] = 0;
            }
            if (a.result) pipeline_result_from_partial(P, a.K, a.result);
            a.ws->counter = 0;
        }
    }
}

__global__ void pipeline_finalize_kernel(const double* __restrict__ partial, int K, double* __restrict__ result) {
    if (blockIdx.x == 0 && threadIdx.x == 0) pipeline_result_from_partial(partial, K, result);
}

struct PipeLaunch {
    PipeArgs a;
    cudaStream_t stream;
    template <int TPM, int NV, int MODE, int MPB>
    void run() const {
        const int grid = (a.n_maps + MPB - 1) / MPB;
        pipeline_kernel<TPM, NV, MODE, MPB><<<grid, TPM * MPB, table_bytes(a.tmp), stream>>>(a);
    }
};

static int launch_pipeline(const float* pred, const double* joints, const float* vis, int B, int K, int H, int W,
                           double stride_x, double stride_y, int tmp, const float* tab, float kl_epsilon, double thr,
                           int loss_mask, float* pred_xy, float* maxvals, float* weight_out, double* partial,
                           int accumulate, double* result, void* workspace, int ws_maps, cudaStream_t stream) {
    PipeArgs a{};
    a.pred = pred; a.joints = joints; a.vis = vis; a.n_maps = B * K; a.K = K; a.H = H; a.W = W;
    a.wdiv = FastDiv(static_cast<uint32_t>(W)); a.sx = stride_x; a.sy = stride_y; a.tmp = tmp; a.tab = tab;
    a.eps = kl_epsilon; a.thr = thr; a.loss_mask = loss_mask; a.pred_xy = pred_xy; a.maxvals = maxvals;
    a.weight_out = weight_out; a.partial = partial; a.accumulate = accumulate; a.result = result;
    a.ws = static_cast<Workspace*>(workspace);
    a.map_vals = reinterpret_cast<float*>(static_cast<char*>(workspace) + (sizeof(Workspace) + 255) / 256 * 256);
    (void)ws_maps;
    const int HW = H * W, side = 2 * tmp + 1;
    // fast shape: one warp per map, register-tile streaming (hp_pipeline_stream.cuh)
    const bool stream_ok = aligned16(pred) && (W % 4 == 0) && (HW % 256 == 0) && HW < (1 << 24) &&
                           side * side <= 32 * kStreamMaxPatch && (sizeof(Workspace) + 255) / 256 * 256 +
                           sizeof(double) * 2 * ((a.n_maps + kStreamWarps - 1) / kStreamWarps) <=
                           hp_workspace_bytes(a.n_maps, K);
    if (stream_ok) {
        StreamArgs sa{};
        sa.pred = pred; sa.joints = joints; sa.vis = vis; sa.n_maps = a.n_maps; sa.K = K; sa.H = H; sa.W = W; sa.HW = HW;
        sa.wdiv = a.wdiv; sa.sdiv = FastDiv(static_cast<uint32_t>(side));
        sa.sx = stride_x; sa.sy = stride_y; sa.inv_sx = 1.0 / stride_x; sa.inv_sy = 1.0 / stride_y;
        int ex = 0, ey = 0;
        sa.pow2_stride = (std::frexp(stride_x, &ex) == 0.5 && std::frexp(stride_y, &ey) == 0.5) ? 1 : 0;
        sa.tmp = tmp; sa.tab = tab; sa.eps = kl_epsilon;
        sa.eps_log_eps = kl_epsilon > 0.0f ? static_cast<float>(static_cast<double>(kl_epsilon) * std::log(static_cast<double>(kl_epsilon))) : 0.0f;
        sa.thr = thr;
        const double t2 = thr * thr;
        sa.thr2_lo = static_cast<float>(t2 * (1.0 - 1e-4)); sa.thr2_hi = static_cast<float>(t2 * (1.0 + 1e-4));
        sa.inv_nx = static_cast<float>(10.0 / H); sa.inv_ny = static_cast<float>(10.0 / W);  // norm = (H/10, W/10) on (x, y)
        sa.pred_xy = pred_xy; sa.maxvals = maxvals; sa.weight_out = weight_out; sa.partial = partial;
        sa.accumulate = accumulate; sa.result = result; sa.ws = a.ws;
        sa.cta_vals = reinterpret_cast<double*>(a.map_vals);
        const int grid = (a.n_maps + kStreamWarps - 1) / kStreamWarps;
        const size_t smem = table_bytes(tmp);
        const bool big = (HW % 1024) == 0;
        sa.ntiles = HW / (big ? 1024 : 256);
#define HP_STREAM_LAUNCH(NVV, LM) pipeline_stream_kernel<NVV, LM><<<grid, 32 * kStreamWarps, smem, stream>>>(sa)
#define HP_STREAM_BY_LOSS(NVV)                                  \
        switch (loss_mask) {                                    \
            case 0: HP_STREAM_LAUNCH(NVV, 0); break;            \
            case HP_LOSS_MSE: HP_STREAM_LAUNCH(NVV, 1); break;  \
            case HP_LOSS_KL: HP_STREAM_LAUNCH(NVV, 2); break;   \
            default: HP_STREAM_LAUNCH(NVV, 3); break;           \
        }
        if (big) { HP_STREAM_BY_LOSS(8) } else { HP_STREAM_BY_LOSS(2) }
#undef HP_STREAM_BY_LOSS
#undef HP_STREAM_LAUNCH
        return launch_status("hp_pipeline_fused");
    }
    PipeLaunch l{a, stream};
    dispatch_map_walk(H * W, aligned16(pred) && (W % 4 == 0), l);
    return launch_status("hp_pipeline_fused");
}

static int check_pipeline(const char* who, const void* pred, const void* joints, const void* vis, int B, int K, int H,
                          int W, double sx, double sy, int tmp, const void* tab, const void* pred_xy,
                          const void* partial, const void* workspace, int loss_mask) {
    HP_REQUIRE(pred && joints && vis && tab && pred_xy && partial && workspace, HP_ERR_NULL, "%s: null pointer", who);
    HP_REQUIRE(B > 0 && K > 0 && K <= HP_MAX_K && H > 0 && W > 0 && static_cast<long long>(H) * W < (1ll << 30) &&
                   static_cast<long long>(B) * K < (1ll << 31),
               HP_ERR_SHAPE, "%s: bad shape B=%d K=%d H=%d W=%d", who, B, K, H, W);
    HP_REQUIRE(tmp >= 0 && tmp <= 64 && sx > 0.0 && sy > 0.0 && (loss_mask & ~(HP_LOSS_MSE | HP_LOSS_KL)) == 0,
               HP_ERR_ARG, "%s: bad tmp/stride/loss_mask", who);
    HP_REQUIRE(aligned8(joints) && aligned8(partial) && aligned8(workspace) && aligned4(pred), HP_ERR_ALIGN,
               "%s: misaligned pointer", who);
    return HP_OK;
}

}  // namespace hp

using namespace hp;

extern "C" HP_API int hp_pipeline_fused(const float* pred, const double* joints, const float* vis, int B, int K, int H,
                                        int W, double stride_x, double stride_y, int tmp, const float* tab,
                                        float kl_epsilon, double thr, int loss_mask, float* pred_xy, float* maxvals,
                                        float* weight_out, double* partial, int accumulate, double* result,
                                        void* workspace, hp_stream_t stream) {
    if (int rc = check_pipeline("hp_pipeline_fused", pred, joints, vis, B, K, H, W, stride_x, stride_y, tmp, tab, pred_xy,
                                partial, workspace, loss_mask))
        return rc;
    return launch_pipeline(pred, joints, vis, B, K, H, W, stride_x, stride_y, tmp, tab, kl_epsilon, thr, loss_mask,
                           pred_xy, maxvals, weight_out, partial, accumulate, result, workspace, B * K,
                           static_cast<cudaStream_t>(stream));
}

extern "C" HP_API int hp_pipeline_finalize(const double* partial, int K, double* result, hp_stream_t stream) {
    HP_REQUIRE(partial && result, HP_ERR_NULL, "hp_pipeline_finalize: null pointer");
    HP_REQUIRE(K > 0 && K <= HP_MAX_K, HP_ERR_SHAPE, "hp_pipeline_finalize: K=%d", K);
    pipeline_finalize_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(partial, K, result);
    return launch_status("hp_pipeline_finalize");
}

extern "C" HP_API int hp_pipeline_fused_host(const float* h_pred, const double* h_joints, const float* h_vis, int B,
                                             int K, int H, int W, double stride_x, double stride_y, int tmp,
                                             const float* tab, float kl_epsilon, double thr, int loss_mask, int slab_B,
                                             float* d_pred, double* d_joints, float* d_vis, float* d_pred_xy,
                                             float* d_maxvals, float* d_weight, double* d_partial, double* d_result,
                                             void* workspace, float* h_pred_xy, double* h_result, hp_stream_t stream,
                                             hp_stream_t copy_stream) {
    if (int rc = check_pipeline("hp_pipeline_fused_host", h_pred, h_joints, h_vis, B, K, H, W, stride_x, stride_y, tmp,
                                tab, d_pred_xy, d_partial, workspace, loss_mask))
        return rc;
    HP_REQUIRE(d_pred && d_joints && d_vis && d_result && h_result, HP_ERR_NULL, "hp_pipeline_fused_host: null pointer");
    HP_REQUIRE(slab_B > 0, HP_ERR_ARG, "hp_pipeline_fused_host: slab_B=%d", slab_B);
    cudaStream_t cs = static_cast<cudaStream_t>(stream), xs = static_cast<cudaStream_t>(copy_stream);
    const size_t map_elems = static_cast<size_t>(H) * W;
    const int n_slabs = (B + slab_B - 1) / slab_B;
    cudaEvent_t filled[2], drained[2];
    cudaError_t e = cudaSuccess;
    for (int i = 0; i < 2 && e == cudaSuccess; ++i) {
        e = cudaEventCreateWithFlags(&filled[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&drained[i], cudaEventDisableTiming);
    }
    if (e != cudaSuccess) return fail(static_cast<int>(e), "hp_pipeline_fused_host: %s", cudaGetErrorString(e));
    int rc = HP_OK;
    // small per-joint inputs go up once; the copy stream must not start before prior work on `stream`
    e = cudaEventRecord(drained[0], cs);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(xs, drained[0], 0);
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(d_joints, h_joints, sizeof(double) * 2 * B * K, cudaMemcpyHostToDevice, xs);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_vis, h_vis, sizeof(float) * B * K, cudaMemcpyHostToDevice, xs);
    for (int s = 0; s < n_slabs && e == cudaSuccess && rc == HP_OK; ++s) {
        const int slot = s & 1, b0 = s * slab_B, nb = (B - b0 < slab_B) ? (B - b0) : slab_B;
        float* slab = d_pred + static_cast<size_t>(slot) * slab_B * K * map_elems;
        if (s >= 2) e = cudaStreamWaitEvent(xs, drained[slot], 0);  // slot's previous kernel is done
        if (e == cudaSuccess)
            e = cudaMemcpyAsync(slab, h_pred + static_cast<size_t>(b0) * K * map_elems,
                                sizeof(float) * nb * K * map_elems, cudaMemcpyHostToDevice, xs);
        if (e == cudaSuccess) e = cudaEventRecord(filled[slot], xs);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(cs, filled[slot], 0);
        if (e != cudaSuccess) break;
        const bool last = (s == n_slabs - 1);
        rc = launch_pipeline(slab, d_joints + 2 * static_cast<size_t>(b0) * K, d_vis + static_cast<size_t>(b0) * K, nb, K,
                             H, W, stride_x, stride_y, tmp, tab, kl_epsilon, thr, loss_mask,
                             d_pred_xy + 2 * static_cast<size_t>(b0) * K,
                             d_maxvals ? d_maxvals + static_cast<size_t>(b0) * K : nullptr,
                             d_weight ? d_weight + static_cast<size_t>(b0) * K : nullptr, d_partial, s > 0 ? 1 : 0,
                             last ? d_result : nullptr, workspace, nb * K, cs);
        if (rc == HP_OK) e = cudaEventRecord(drained[slot], cs);
    }
    if (e == cudaSuccess && rc == HP_OK)
        e = cudaMemcpyAsync(h_result, d_result, sizeof(double) * (4 + K), cudaMemcpyDeviceToHost, cs);
    if (e == cudaSuccess && rc == HP_OK && h_pred_xy)
        e = cudaMemcpyAsync(h_pred_xy, d_pred_xy, sizeof(float) * 2 * B * K, cudaMemcpyDeviceToHost, cs);
    cudaError_t e2 = cudaStreamSynchronize(cs);
    cudaStreamSynchronize(xs);
    for (int i = 0; i < 2; ++i) {
        cudaEventDestroy(filled[i]);
        cudaEventDestroy(drained[i]);
    }
    if (rc != HP_OK) return rc;
    if (e == cudaSuccess) e = e2;
    if (e != cudaSuccess) return fail(static_cast<int>(e), "hp_pipeline_fused_host: %s", cudaGetErrorString(e));
    return HP_OK;
}
