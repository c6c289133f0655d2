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
#include <atomic>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <new>

#include "hp_common.cuh"
#include "hp_dispatch.cuh"
#include "hp_internal.cuh"
#include "hp_pipeline_common.cuh"
#include "hp_pipeline_parts.cuh"
#include "hp_pipeline_bulk.cuh"
#include "hp_pipeline_stream.cuh"

namespace hp {

// ---------------------------------------------------------------------------------------------
// generic shape: any H, W (odd widths, unaligned bases), any patch size.  Block-per-map register
// tiles with the target evaluated next to every element - correct everywhere, fast nowhere; the
// aligned power-of-two shapes never come here.
// ---------------------------------------------------------------------------------------------
template <int TPM, int NV, int MODE, int MPB>
__global__ void __launch_bounds__(TPM* MPB) pipeline_generic_kernel(const PipeArgs a) {
    // per-map sums: 0 sum (p-t)^2 | 1 sum p | 2 sum_patch u*p | 3 sum_patch u*log(u) | 4 sum_patch u | 5 sum_patch p
    constexpr int NS = 6;
    extern __shared__ float s_tab[];
    __shared__ Stats<NS> scratch[TPM > 32 ? TPM / 32 + 1 : 1];
    __shared__ Centre s_centre[MPB];
    __shared__ float s_weight[MPB];
    __shared__ double s_map[2][MPB];
    __shared__ BlockLoss s_loss;

    const int HW = a.HW;
    const int g = threadIdx.x / TPM, t = threadIdx.x % TPM;
    const int map = blockIdx.x * MPB + g;
    const bool active = map < a.n_maps;
    const bool want_mse = (a.loss_mask & HP_LOSS_MSE) != 0, want_kl = (a.loss_mask & HP_LOSS_KL) != 0;
    const float* pm = a.pred + static_cast<size_t>(active ? map : 0) * HW;
    const int ntiles = (MODE == WALK_EXACT) ? 1 : tiles_for<TPM, NV>(HW);

    float4 p[NV];
    if (active) load_tile<TPM, NV, MODE>(pm, HW, 0, t, -INFINITY, p);
    load_table(s_tab, a.tab, a.tmp);
    if (threadIdx.x == 0) block_loss_zero(&s_loss);
    if (t == 0) s_map[0][g] = s_map[1][g] = 0.0;
    if (active && t == 0) {
        float w;
        s_centre[g] = pipe_centre(a, a.joints[2 * map], a.joints[2 * map + 1], a.vis[map], w);
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
            double mse = 0.0, kl = 0.0;
            if (want_mse)  // mean over HW of 0.5*w*(p-t)^2   (loss.py:59-65)
                mse = 0.5 * static_cast<double>(w) * static_cast<double>(st.sum[0]) / static_cast<double>(HW);
            if (want_kl) {
                const double eps = static_cast<double>(a.eps);
                const double n_bg = static_cast<double>(HW - pipe_patch_area(a, c));
                const double Su = static_cast<double>(st.sum[4]) + eps * n_bg;
                const double Sup = static_cast<double>(st.sum[2]) +
                                   eps * (static_cast<double>(st.sum[1]) - static_cast<double>(st.sum[5]));
                const double Sulogu = static_cast<double>(st.sum[3]) + ((a.eps > 0.0f) ? n_bg * eps * log(eps) : 0.0);
                const double lse = static_cast<double>(st.m) + log(static_cast<double>(st.s));
                const double L = (Sulogu - Sup) / Su - log(Su) + lse;  // Su == 0 (eps 0, nothing pasted) -> NaN
                kl = L * static_cast<double>(w);
            }
            s_map[0][g] = mse;
            s_map[1][g] = kl;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int n_here = min(MPB, a.n_maps - static_cast<int>(blockIdx.x) * MPB);
        for (int w = 0; w < n_here; ++w) {
            if (want_mse) block_loss_add(&s_loss, 0, s_map[0][w]);
            if (want_kl) block_loss_add(&s_loss, 1, s_map[1][w]);
        }
        block_loss_flush(&s_loss, a.ws);
    }
    if (pipeline_last_block(a.ws)) pipeline_publish(a);
}

__global__ void pipeline_finalize_kernel(const long long* __restrict__ partial, int K, double* __restrict__ result) {
    if (blockIdx.x == 0 && threadIdx.x == 0) pipeline_result_from_partial(partial, K, result);
}

struct PipeLaunch {
    PipeArgs a;
    cudaStream_t stream;
    template <int TPM, int NV, int MODE, int MPB>
    void run() const {
        const int grid = (a.n_maps + MPB - 1) / MPB;
        pipeline_generic_kernel<TPM, NV, MODE, MPB><<<grid, TPM * MPB, table_bytes(a.tmp), stream>>>(a);
    }
};

static int g_sm_count = 0;
// profiling only (hp_debug_pipeline_trace): two launch slots, alternating
static unsigned long long* g_trace = nullptr;
static size_t g_trace_words = 0;
static unsigned long long g_trace_seq = 0;

// experiments: HP_PIPE_GRID overrides the number of blocks of the TMA-staged kernel
static int pipeline_grid_override() {
    static const int v = []() {
        const char* e = std::getenv("HP_PIPE_GRID");
        const int g = e ? std::atoi(e) : 0;
        return g > 0 ? g : 0;
    }();
    return v;
}
// experiments: HP_PIPE_GRID_DIV overrides the depth requested through the flags (0 = no override)
static int pipeline_grid_div_override() {
    static const int div = []() {
        const char* e = std::getenv("HP_PIPE_GRID_DIV");
        const int v = e ? std::atoi(e) : 0;
        return (v >= 1 && v <= 8) ? v : 0;
    }();
    return div;
}
// ---- TMA-staged shape (hp_pipeline_bulk.cuh) ---------------------------------------------------------------------
// experiments: HP_PIPE_EPILOGUE=atomics keeps the workspace-atomics + ticket epilogue; HP_PIPE_STRICT_PDL=0 launches a
// serialised step without the programmatic attribute (the next launch is then not even scheduled before this one ends)
// HP_PIPE_EPILOGUE=atomics keeps the workspace-atomics + fence + ticket epilogue (comparison runs)
static bool pipeline_cert_enabled() {
    static const bool on = []() {
        const char* e = std::getenv("HP_PIPE_EPILOGUE");
        return !(e && (e[0] == 'a' || e[0] == 'A'));
    }();
    return on;
}
static bool pipeline_strict_pdl() {
    static const bool on = []() {
        const char* e = std::getenv("HP_PIPE_STRICT_PDL");
        return !(e && e[0] == '0');
    }();
    return on;
}

template <int NITC, int LOSS, bool MULTI, int W, int KST, int BPS>
static cudaError_t launch_bulk_one(BulkArgs& t, int grid, cudaStream_t stream) {
    constexpr size_t smem = static_cast<size_t>(W) * KST * NITC * 512 + sizeof(uint64_t) * W * KST;
    static bool configured[16] = {};  // per device: opt in to > 48 KB of dynamic shared memory once
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 16 || !configured[dev]) {
        const cudaError_t e = cudaFuncSetAttribute(pipeline_bulk_kernel<NITC, LOSS, MULTI, W, KST, BPS>,
                                                   cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 16) configured[dev] = true;
    }
    // ---- how the blocks hand their sums to the publisher (hp_pipeline_bulk.cuh "self-certifying accumulators") --------
    t.certs = reinterpret_cast<unsigned long long*>(reinterpret_cast<unsigned char*>(t.p.ws) + kCertOffsetBytes);
    t.cert = (pipeline_cert_enabled() && grid <= 65535 && t.p.n_maps <= 65535) ? 1 : 0;  // 16-bit count fields
    t.strict = (t.overlap == 0 && pipeline_strict_pdl()) ? 1 : 0;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(static_cast<unsigned>(grid));
    cfg.blockDim = dim3(32 * W);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (t.overlap > 0 || t.strict) ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, pipeline_bulk_kernel<NITC, LOSS, MULTI, W, KST, BPS>, t);
}
template <int NITC, bool MULTI, int W, int KST, int BPS>
static cudaError_t launch_bulk(BulkArgs& t, int sms, cudaStream_t stream) {
    // persistent: BPS blocks per SM.  An overlapped launch may take only 1/div of the slots so that `div`
    // consecutive launches are resident at once, out of phase (their start-up and drain bubbles interleave).
    // The depth only pays while a launch is a few rounds long (its start-up and drain are then a third of its
    // life); it is capped so that a block's per-map outputs still fit its shared-memory buffer - beyond that a
    // block has to wait for the previous launch in mid-stream, which serialises the train on 1/depth of the SMs.
    int slots = sms * BPS;
    if (t.overlap > 1) {
        const long long fit = static_cast<long long>(kBulkOutCap) * slots / (t.p.n_maps > 0 ? t.p.n_maps : 1);
        const int depth = fit < 1 ? 1 : (fit < t.overlap ? static_cast<int>(fit) : t.overlap);
        slots = (slots + depth - 1) / depth;
    }
    // (a smaller grid that gives every warp the same number of maps - 112 x 12 x 4 rounds for 256 x 21 maps - was measured
    // SLOWER: with fewer blocks the copy engines no longer saturate HBM, profiles/r2_pipeline_history.md)
    int grid = t.p.n_maps < slots ? t.p.n_maps : slots;
    if (const int g = pipeline_grid_override()) grid = g < t.p.n_maps ? g : t.p.n_maps;
    switch (t.p.loss_mask) {
        case 0: return launch_bulk_one<NITC, 0, MULTI, W, KST, BPS>(t, grid, stream);
        case HP_LOSS_MSE: return launch_bulk_one<NITC, 1, MULTI, W, KST, BPS>(t, grid, stream);
        case HP_LOSS_KL: return launch_bulk_one<NITC, 2, MULTI, W, KST, BPS>(t, grid, stream);
        default: return launch_bulk_one<NITC, 3, MULTI, W, KST, BPS>(t, grid, stream);
    }
}
// experiments: HP_PIPE_SHAPE = 0 / "stream": never the TMA-staged kernel; 3: one block of 12 warps per SM for 64x64 maps
static bool pipeline_shape_forced() { return std::getenv("HP_PIPE_SHAPE") != nullptr; }
static int pipeline_shape_choice() {
    static const int choice = []() {
        const char* e = std::getenv("HP_PIPE_SHAPE");
        if (!e) return 1;
        if (e[0] == '0' || e[0] == 's' || e[0] == 'S') return 0;
        if (e[0] == '3') return 3;
        return 1;
    }();
    return choice;
}

static int launch_pipeline(const float* pred, const double* joints, const float* vis, int B, int K, int H, int W,
                           double stride_x, double stride_y, int tmp, const float* tab, float kl_epsilon, double thr,
                           int loss_mask, float* pred_xy, float* maxvals, float* weight_out, long long* partial,
                           int accumulate, double* result, void* workspace, cudaStream_t stream, unsigned flags = 0,
                           const PeerLink* link = nullptr, bool* exchanged = nullptr) {
    if (exchanged) *exchanged = false;
    const int HW = H * W, side = 2 * tmp + 1;
    PipeArgs a{};
    a.pred = pred; a.joints = joints; a.vis = vis; a.n_maps = B * K; a.K = K; a.H = H; a.W = W; a.HW = HW;
    a.wdiv = FastDiv(static_cast<uint32_t>(W)); a.sdiv = FastDiv(static_cast<uint32_t>(side));
    a.sx = stride_x; a.sy = stride_y; a.inv_sx = 1.0 / stride_x; a.inv_sy = 1.0 / stride_y;
    int ex = 0, ey = 0;
    a.pow2_stride = (std::frexp(stride_x, &ex) == 0.5 && std::frexp(stride_y, &ey) == 0.5) ? 1 : 0;
    a.tmp = tmp; a.tab = tab; a.eps = kl_epsilon;
    a.eps_log_eps = kl_epsilon > 0.0f
                        ? static_cast<float>(static_cast<double>(kl_epsilon) * std::log(static_cast<double>(kl_epsilon)))
                        : 0.0f;
    a.thr = thr;
    const double t2 = thr * thr;
    a.thr2_lo = static_cast<float>(t2 * (1.0 - 1e-4)); a.thr2_hi = static_cast<float>(t2 * (1.0 + 1e-4));
    if (thr <= 0.0) a.thr2_lo = a.thr2_hi = -1.0f;  // d < thr never holds (the fp32 pre-test must not see thr^2)
    a.inv_nx = static_cast<float>(10.0 / H); a.inv_ny = static_cast<float>(10.0 / W);  // norm = (H/10, W/10) on (x, y)
    a.loss_mask = loss_mask; a.pred_xy = pred_xy; a.maxvals = maxvals; a.weight_out = weight_out;
    a.partial = partial; a.accumulate = accumulate; a.result = result; a.ws = static_cast<Workspace*>(workspace);

    const bool fast_ok = aligned16(pred) && (W % 4 == 0) && HW < (1 << 24) && side * side <= 32 * kTileMaxPatch &&
                         thr > 0.0;
    const size_t smem = table_bytes(tmp);  // the generic / stream shapes stage the Gaussian table here
#define HP_BY_LOSS(KERNEL, GRID, ARG, NVV)                                                       \
    switch (loss_mask) {                                                                         \
        case 0: KERNEL<NVV, 0><<<GRID, 128, smem, stream>>>(ARG); break;                         \
        case HP_LOSS_MSE: KERNEL<NVV, 1><<<GRID, 128, smem, stream>>>(ARG); break;               \
        case HP_LOSS_KL: KERNEL<NVV, 2><<<GRID, 128, smem, stream>>>(ARG); break;                \
        default: KERNEL<NVV, 3><<<GRID, 128, smem, stream>>>(ARG); break;                        \
    }
    if (fast_ok && pipeline_shape_choice() != 0 &&
        (HW == 256 || HW == 1024 || (HW % 4096 == 0 && HW / 4096 <= 64))) {
        if (g_sm_count == 0) {
            g_sm_count = hp_device_sm_count();
            if (g_sm_count <= 0) g_sm_count = 148;
        }
        BulkArgs t{};
        t.p = a;
        t.kdiv = FastDiv(static_cast<uint32_t>(K));
        t.n_chunks = HW > 4096 ? HW / 4096 : 1;
        if (flags & HP_PIPE_OVERLAP_PREV) {  // depth of the launch train: 0 -> library default
            const int depth = static_cast<int>((flags >> 8) & 15u);
            t.overlap = pipeline_grid_div_override() ? pipeline_grid_div_override() : (depth ? depth : 4);
        }
        if (link && link->world > 1 && accumulate == 0) {  // the kernel's last block does the exchange itself
            HP_REQUIRE(peer_shape_ok(K, link->world), HP_ERR_SHAPE,
                       "hp_pipeline_fused: K=%d world=%d exceeds the peer mailbox (2*(4+2K+6) <= %d entries per source, "
                       "world*(4+2K+6) <= 1024)", K, link->world, kPeerSlotEntries);
            t.link = *link;
            t.defer = (flags & HP_PIPE_DEFER_EXCHANGE) ? 1 : 0;
            if (exchanged) *exchanged = true;
        }
        const size_t trace_words = static_cast<size_t>(kTraceBlockWords) * kTraceMaxBlocksPerSM * static_cast<size_t>(g_sm_count);
        if (g_trace && g_trace_words >= 2 * trace_words) t.trace = g_trace + (g_trace_seq++ & 1) * trace_words;
        const int sms = g_sm_count;
        cudaError_t e;
        // default: 3 blocks of 4 warps per SM, one 16 KB stage per warp (12 stages = 192 KB in flight per SM)
        if (HW == 256) e = launch_bulk<2, false, 4, 8, 3>(t, sms, stream);
        else if (HW == 1024) e = launch_bulk<8, false, 4, 4, 3>(t, sms, stream);
        else if (HW > 4096) e = launch_bulk<32, true, 4, 1, 3>(t, sms, stream);
        // 64x64: a train of launches takes 3 blocks of 4 warps per SM (the next launch takes an SM over block by block);
        // a serialised launch one block of 12 warps per SM - the same 12 stages in flight, a third of the blocks in the
        // epilogue (21.4 vs 22.3-24 us per launch, profiles/r2_pipeline_history.md).  HP_PIPE_SHAPE=1 / 3 force one.
        else if (pipeline_shape_choice() == 3 || (pipeline_shape_choice() == 1 && t.overlap == 0 && !pipeline_shape_forced()))
            e = launch_bulk<32, false, 12, 1, 1>(t, sms, stream);
        else e = launch_bulk<32, false, 4, 1, 3>(t, sms, stream);
        if (e != cudaSuccess) return fail(static_cast<int>(e), "hp_pipeline_fused: %s", cudaGetErrorString(e));
        return launch_status("hp_pipeline_fused");
    }
    const int tile_elems = (HW % 1024 == 0) ? 1024 : ((HW % 256 == 0) ? 256 : 0);
    if (fast_ok && tile_elems != 0) {
        // aligned maps of whole 1 KB / 4 KB tiles that the TMA-staged kernel does not take (e.g. 32x64, 48x64): one warp
        // streams a whole map through register tiles
        const int grid = (a.n_maps + kStreamWarps - 1) / kStreamWarps;
        a.ntiles = HW / tile_elems;
        if (tile_elems == 1024) { HP_BY_LOSS(pipeline_stream_kernel, grid, a, 8) } else { HP_BY_LOSS(pipeline_stream_kernel, grid, a, 2) }
        return launch_status("hp_pipeline_fused");
    }
#undef HP_BY_LOSS
    PipeLaunch l{a, stream};
    dispatch_map_walk(HW, aligned16(pred) && (W % 4 == 0), l);
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
                                        float* weight_out, int64_t* partial, int accumulate, double* result,
                                        void* workspace, hp_stream_t stream) {
    if (int rc = check_pipeline("hp_pipeline_fused", pred, joints, vis, B, K, H, W, stride_x, stride_y, tmp, tab, pred_xy,
                                partial, workspace, loss_mask))
        return rc;
    return launch_pipeline(pred, joints, vis, B, K, H, W, stride_x, stride_y, tmp, tab, kl_epsilon, thr, loss_mask,
                           pred_xy, maxvals, weight_out, reinterpret_cast<long long*>(partial), accumulate, result,
                           workspace, static_cast<cudaStream_t>(stream));
}

extern "C" HP_API int hp_pipeline_fused_ex(const float* pred, const double* joints, const float* vis, int B, int K, int H,
                                           int W, double stride_x, double stride_y, int tmp, const float* tab,
                                           float kl_epsilon, double thr, int loss_mask, float* pred_xy, float* maxvals,
                                           float* weight_out, int64_t* partial, int accumulate, double* result,
                                           void* workspace, unsigned int flags, hp_stream_t stream) {
    if (int rc = check_pipeline("hp_pipeline_fused_ex", pred, joints, vis, B, K, H, W, stride_x, stride_y, tmp, tab,
                                pred_xy, partial, workspace, loss_mask))
        return rc;
    HP_REQUIRE((flags & ~(HP_PIPE_OVERLAP_PREV | HP_PIPE_DEPTH(15u))) == 0 && ((flags >> 8) & 15u) <= 8u, HP_ERR_ARG,
               "hp_pipeline_fused_ex: bad flags 0x%x", flags);
    return launch_pipeline(pred, joints, vis, B, K, H, W, stride_x, stride_y, tmp, tab, kl_epsilon, thr, loss_mask,
                           pred_xy, maxvals, weight_out, reinterpret_cast<long long*>(partial), accumulate, result,
                           workspace, static_cast<cudaStream_t>(stream), flags);
}

/* one sharded step: the fused kernel on this rank's slice, then the exchange + finalise kernel, both on `stream`
 * (with HP_PIPE_OVERLAP_PREV both are programmatic dependent launches: a train of steps stays overlapped) */
extern "C" HP_API int hp_pipeline_fused_peer(const float* pred, const double* joints, const float* vis, int B, int K,
                                             int H, int W, double stride_x, double stride_y, int tmp, const float* tab,
                                             float kl_epsilon, double thr, int loss_mask, float* pred_xy,
                                             float* maxvals, float* weight_out, int64_t* partial, double* result,
                                             void* workspace, void* const* mailboxes, int rank, int world,
                                             int64_t seq, unsigned int flags, hp_stream_t stream) {
    if (int rc = check_pipeline("hp_pipeline_fused_peer", pred, joints, vis, B, K, H, W, stride_x, stride_y, tmp, tab,
                                pred_xy, partial, workspace, loss_mask))
        return rc;
    HP_REQUIRE((flags & ~(HP_PIPE_OVERLAP_PREV | HP_PIPE_DEFER_EXCHANGE | HP_PIPE_DEPTH(15u))) == 0 && ((flags >> 8) & 15u) <= 8u,
               HP_ERR_ARG, "hp_pipeline_fused_peer: bad flags 0x%x", flags);
    HP_REQUIRE(result && mailboxes, HP_ERR_NULL, "hp_pipeline_fused_peer: null pointer");
    HP_REQUIRE(world > 0 && world <= kPeerMaxWorld && rank >= 0 && rank < world && seq == 0, HP_ERR_ARG,
               "hp_pipeline_fused_peer: rank=%d world=%d seq=%lld (the step is counted on the device: pass 0)", rank,
               world, static_cast<long long>(seq));
    HP_REQUIRE(peer_shape_ok(K, world), HP_ERR_SHAPE,
               "hp_pipeline_fused_peer: K=%d world=%d exceeds the peer mailbox (K <= 27 when sharded)", K, world);
    PeerLink link{};
    for (int r = 0; r < world; ++r) {
        HP_REQUIRE(mailboxes[r], HP_ERR_NULL, "hp_pipeline_fused_peer: mailbox %d is null", r);
        link.mailbox[r] = static_cast<unsigned long long*>(mailboxes[r]);
    }
    link.rank = rank;
    link.world = world;
    bool exchanged = false;
    // shapes served by the TMA-staged kernel: its last block exchanges and finalises in place (ONE kernel per step)
    if (int rc = launch_pipeline(pred, joints, vis, B, K, H, W, stride_x, stride_y, tmp, tab, kl_epsilon, thr, loss_mask,
                                 pred_xy, maxvals, weight_out, reinterpret_cast<long long*>(partial), 0, result,
                                 workspace, static_cast<cudaStream_t>(stream), flags, &link, &exchanged))
        return rc;
    if (exchanged || world == 1) return HP_OK;
    // other shapes: the separate exchange + finalise kernel
    return launch_finalize_peer(reinterpret_cast<const long long*>(partial), mailboxes, rank, world, K, 0,
                                reinterpret_cast<long long*>(partial), result,
                                (flags & HP_PIPE_OVERLAP_PREV) ? 1 : 0, workspace, (flags & HP_PIPE_DEFER_EXCHANGE) ? 1 : 0,
                                static_cast<cudaStream_t>(stream));
}

/* ---- pre-bound steps -------------------------------------------------------------------------------------------
 * A step of the pipeline is a ~12 us kernel; marshalling two dozen arguments through an FFI per step costs about
 * as much on the host.  A plan validates and stores the arguments of hp_pipeline_fused_ex / hp_pipeline_fused_peer
 * once; launching it is a two-argument call.  The plan holds the caller's pointers until it is destroyed. */
struct hp_plan {
    const float* pred; const double* joints; const float* vis;
    int B, K, H, W; double sx, sy; int tmp; const float* tab; float eps; double thr; int loss_mask;
    float* pred_xy; float* maxvals; float* weight_out; long long* partial; int accumulate; double* result;
    void* workspace; unsigned flags;
    PeerLink link;                       // world <= 1: single rank
    void* mailboxes[kPeerMaxWorld];
};

extern "C" HP_API int hp_pipeline_plan_create(const float* pred, const double* joints, const float* vis, int B, int K,
                                              int H, int W, double stride_x, double stride_y, int tmp, const float* tab,
                                              float kl_epsilon, double thr, int loss_mask, float* pred_xy,
                                              float* maxvals, float* weight_out, int64_t* partial, int accumulate,
                                              double* result, void* workspace, void* const* mailboxes, int rank,
                                              int world, unsigned int flags, hp_plan** plan) {
    HP_REQUIRE(plan, HP_ERR_NULL, "hp_pipeline_plan_create: null plan pointer");
    *plan = nullptr;
    if (int rc = check_pipeline("hp_pipeline_plan_create", pred, joints, vis, B, K, H, W, stride_x, stride_y, tmp, tab,
                                pred_xy, partial, workspace, loss_mask))
        return rc;
    HP_REQUIRE((flags & ~(HP_PIPE_OVERLAP_PREV | HP_PIPE_DEFER_EXCHANGE | HP_PIPE_DEPTH(15u))) == 0 && ((flags >> 8) & 15u) <= 8u,
               HP_ERR_ARG, "hp_pipeline_plan_create: bad flags 0x%x", flags);
    HP_REQUIRE(!(flags & HP_PIPE_DEFER_EXCHANGE) || world > 1, HP_ERR_ARG,
               "hp_pipeline_plan_create: HP_PIPE_DEFER_EXCHANGE needs a sharded step (world > 1)");
    HP_REQUIRE(world >= 0 && world <= kPeerMaxWorld && (world <= 1 || (mailboxes && rank >= 0 && rank < world && result &&
                                                                      accumulate == 0)),
               HP_ERR_ARG, "hp_pipeline_plan_create: rank=%d world=%d", rank, world);
    HP_REQUIRE(peer_shape_ok(K, world), HP_ERR_SHAPE,
               "hp_pipeline_plan_create: K=%d world=%d exceeds the peer mailbox (K <= 27 when sharded)", K, world);
    hp_plan* p = new (std::nothrow) hp_plan{};
    HP_REQUIRE(p, HP_ERR_ARG, "hp_pipeline_plan_create: out of host memory");
    p->pred = pred; p->joints = joints; p->vis = vis; p->B = B; p->K = K; p->H = H; p->W = W;
    p->sx = stride_x; p->sy = stride_y; p->tmp = tmp; p->tab = tab; p->eps = kl_epsilon; p->thr = thr;
    p->loss_mask = loss_mask; p->pred_xy = pred_xy; p->maxvals = maxvals; p->weight_out = weight_out;
    p->partial = reinterpret_cast<long long*>(partial); p->accumulate = accumulate; p->result = result;
    p->workspace = workspace; p->flags = flags;
    p->link.rank = rank; p->link.world = world > 1 ? world : 0;
    for (int r = 0; r < p->link.world; ++r) {
        if (!mailboxes[r]) {
            delete p;
            return fail(HP_ERR_NULL, "hp_pipeline_plan_create: mailbox %d is null", r);
        }
        p->mailboxes[r] = mailboxes[r];
        p->link.mailbox[r] = static_cast<unsigned long long*>(mailboxes[r]);
    }
    *plan = p;
    return HP_OK;
}

extern "C" HP_API int hp_pipeline_plan_launch(const hp_plan* p, hp_stream_t stream) {
    HP_REQUIRE(p, HP_ERR_NULL, "hp_pipeline_plan_launch: null plan");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (p->link.world <= 1)
        return launch_pipeline(p->pred, p->joints, p->vis, p->B, p->K, p->H, p->W, p->sx, p->sy, p->tmp, p->tab, p->eps,
                               p->thr, p->loss_mask, p->pred_xy, p->maxvals, p->weight_out, p->partial, p->accumulate,
                               p->result, p->workspace, st, p->flags);
    bool exchanged = false;
    if (int rc = launch_pipeline(p->pred, p->joints, p->vis, p->B, p->K, p->H, p->W, p->sx, p->sy, p->tmp, p->tab, p->eps,
                                 p->thr, p->loss_mask, p->pred_xy, p->maxvals, p->weight_out, p->partial, 0, p->result,
                                 p->workspace, st, p->flags, &p->link, &exchanged))
        return rc;
    if (exchanged) return HP_OK;
    return launch_finalize_peer(p->partial, p->mailboxes, p->link.rank, p->link.world, p->K, 0, p->partial, p->result,
                                (p->flags & HP_PIPE_OVERLAP_PREV) ? 1 : 0, p->workspace,
                                (p->flags & HP_PIPE_DEFER_EXCHANGE) ? 1 : 0, st);
}

extern "C" HP_API int hp_pipeline_plan_destroy(hp_plan* p) {
    delete p;
    return HP_OK;
}

/* profiling aid: blocks of the TMA-staged pipeline kernel stamp their timeline into `buf` (device memory,
 * uint64 words; two alternating launch slots of hp_debug_pipeline_trace_words() words each).  NULL switches it off. */
extern "C" HP_API size_t hp_debug_pipeline_trace_words(void) {
    if (g_sm_count == 0) {
        g_sm_count = hp_device_sm_count();
        if (g_sm_count <= 0) g_sm_count = 148;
    }
    return static_cast<size_t>(kTraceBlockWords) * kTraceMaxBlocksPerSM * static_cast<size_t>(g_sm_count);
}
extern "C" HP_API int hp_debug_pipeline_trace(void* buf, size_t words) {
    g_trace = static_cast<unsigned long long*>(buf);
    g_trace_words = buf ? words : 0;
    g_trace_seq = 0;
    return HP_OK;
}

extern "C" HP_API int hp_pipeline_finalize(const int64_t* partial, int K, double* result, hp_stream_t stream) {
    HP_REQUIRE(partial && result, HP_ERR_NULL, "hp_pipeline_finalize: null pointer");
    HP_REQUIRE(K > 0 && K <= HP_MAX_K, HP_ERR_SHAPE, "hp_pipeline_finalize: K=%d", K);
    pipeline_finalize_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const long long*>(partial),
                                                                              K, result);
    return launch_status("hp_pipeline_finalize");
}

extern "C" HP_API int hp_pipeline_fused_host(const float* h_pred, const double* h_joints, const float* h_vis, int B,
                                             int K, int H, int W, double stride_x, double stride_y, int tmp,
                                             const float* tab, float kl_epsilon, double thr, int loss_mask, int slab_B,
                                             float* d_pred, double* d_joints, float* d_vis, float* d_pred_xy,
                                             float* d_maxvals, float* d_weight, int64_t* d_partial, double* d_result,
                                             void* workspace, float* h_pred_xy, double* h_result,
                                             void* const* mailboxes, int rank, int world, hp_stream_t stream,
                                             hp_stream_t copy_stream) {
    if (int rc = check_pipeline("hp_pipeline_fused_host", h_pred, h_joints, h_vis, B, K, H, W, stride_x, stride_y, tmp,
                                tab, d_pred_xy, d_partial, workspace, loss_mask))
        return rc;
    HP_REQUIRE(d_pred && d_joints && d_vis && d_result && h_result, HP_ERR_NULL, "hp_pipeline_fused_host: null pointer");
    HP_REQUIRE(slab_B > 0, HP_ERR_ARG, "hp_pipeline_fused_host: slab_B=%d", slab_B);
    const bool sharded = world > 1;
    HP_REQUIRE(!sharded || (mailboxes && rank >= 0 && rank < world && world <= kPeerMaxWorld && peer_shape_ok(K, world)),
               HP_ERR_ARG, "hp_pipeline_fused_host: rank=%d world=%d K=%d (K <= 27 when sharded)", rank, world, K);
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
                             d_weight ? d_weight + static_cast<size_t>(b0) * K : nullptr,
                             reinterpret_cast<long long*>(d_partial), s > 0 ? 1 : 0, (last && !sharded) ? d_result : nullptr,
                             workspace, cs);
        if (rc == HP_OK) e = cudaEventRecord(drained[slot], cs);
    }
    if (e == cudaSuccess && rc == HP_OK && sharded)  // the path's one collective: this rank's totals -> totals over the ranks
        rc = launch_finalize_peer(reinterpret_cast<const long long*>(d_partial), mailboxes, rank, world, K, 0,
                                  reinterpret_cast<long long*>(d_partial), d_result, 0, nullptr, 0, cs);
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
