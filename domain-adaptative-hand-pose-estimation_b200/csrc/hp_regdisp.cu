// hp_regdisp.cu - pseudo labels (a6/a7) and the regression-disparity term (a8-a11).
//
// Replaces  PseudoLabelGenerator{,01,03}.forward   uda/model/regda_4.py:76-86, regda_7.py:3026-3039, 3188-3201
//           RegressionDisparity{,x1,x5,x6}.forward uda/model/regda_4.py:129-143, regda_7.py:3250-3268,
//                                                  3529-3561, 3609-3632   (criterion = JointsKLLoss(eps))
// The reference copies y to the host, decodes it with numpy, gathers from a 64 MiB table of
// Gaussians, runs an sgemm for the ground-false map, copies two full tensors back and then
// normalises each map with 2*B*K tiny launches.  Here:
//   launch 1  decode y -> integer centres [B*K,2]                    (hp_decode.cu, reads y once)
//   launch 2  per (sample, joint range) block: rebuild gt / gf in registers from the K centres
//             (sum over joints kept in shared memory), stream y_adv (+ the fused map) once and
//             accumulate the KL statistics; gt / gf are never written unless asked for.
// Ground-false recipes, all fp32, clip after every step:
//   base  clip(sum_{j!=k} gt_j)              x1   clip(1 - 10 gt)
//   x5    g = clip(1 - 10 gt);               [fused f given]  g = clip((g + f) - 100 gt);  g / max(g)
//   x6    g = clip(clip(sum_j gt_j) - 10 gt); [fused f given] g = clip((g + f) - 100 gt);  g / max(g)
// Roofline: HBM.  Bytes per map: y (H*W*4) + y_adv (oh*ow*4) [+ fused oh*ow*4]; backward adds a write.
#include <cmath>
#include <cstdlib>

#include "hp_common.cuh"
#include "hp_internal.cuh"
#include "hp_regdisp_staged.cuh"
#include "hp_regdisp_min.cuh"
#include "hp_regdisp_dense.cuh"
#include "hp_regdisp_sparse.cuh"

namespace hp {

constexpr int kRDThreads = 256;
constexpr int kRDNV = 4;

// un-normalised ground-false value of joint k at pixel (x, y) / flat idx
__device__ __forceinline__ float ground_false_value(int variant, bool has_fused, int k, int K, int x, int y, int idx,
                                                    float gt, float f, const float* s_tab, int tmp, const Centre* s_c,
                                                    const float* s_all) {
    float g;
    if (variant == HP_RD_BASE) {
        float sum;
        if (gt != 0.0f) {  // inside joint k's own patch: exclude it explicitly (no cancellation)
            sum = 0.0f;
            for (int j = 0; j < K; ++j)
                if (j != k) sum += patch_at(s_tab, tmp, s_c[j], x, y);
        } else {
            sum = s_all[idx];
        }
        return clip01(sum);
    }
    if (variant == HP_RD_X6 || variant == HP_RD_RD4) g = clip01(__fsub_rn(clip01(s_all[idx]), __fmul_rn(gt, 10.0f)));
    else g = clip01(__fsub_rn(1.0f, __fmul_rn(gt, 10.0f)));
    if (has_fused && variant != HP_RD_X1 && variant != HP_RD_RD4) g = clip01(__fsub_rn(__fadd_rn(g, f), __fmul_rn(gt, 100.0f)));
    return g;
}

template <int MODE, int TASK>
__global__ void __launch_bounds__(kRDThreads) regdisp_kernel(const RDArgs a) {
    extern __shared__ float s_rdg[];
    __shared__ Centre s_c[HP_MAX_K];
    __shared__ Stats<3> scratch[kRDThreads / 32 + 1];
    __shared__ double s_red[kRDThreads];

    const int ohw = a.oh * a.ow;
    const int ntab = 2 * a.tmp * a.tmp + 1;
    float* s_tab = s_rdg;
    float* s_all = s_rdg + ((ntab + 3) & ~3);
    const int b = blockIdx.x / a.splits, part = blockIdx.x - b * a.splits;
    const int k_begin = (part * a.K) / a.splits, k_end = ((part + 1) * a.K) / a.splits;
    const int t = threadIdx.x;
    const bool want_gf = (TASK == RD_MATERIALIZE) || (a.mode == HP_MODE_MAX);
    const bool needs_all = want_gf && (a.variant == HP_RD_BASE || a.variant == HP_RD_X6 || a.variant == HP_RD_RD4);
    const bool normalise = want_gf && (a.variant == HP_RD_X5 || a.variant == HP_RD_X6);
    const bool has_fused = a.fused != nullptr;

    load_table(s_tab, a.tab, a.tmp);
    if (t < a.K) {
        Centre c;
        c.x = a.centres[2 * (b * a.K + t) + 0];
        c.y = a.centres[2 * (b * a.K + t) + 1];
        s_c[t] = c;
    }
    __syncthreads();
    if (needs_all) {  // sum over all joints of their Gaussians, ascending joint order
        for (int idx = t; idx < ohw; idx += kRDThreads) {
            uint32_t y, x;
            a.wdiv.divmod(static_cast<uint32_t>(idx), y, x);
            float sum = 0.0f;
            for (int j = 0; j < a.K; ++j) sum += patch_at(s_tab, a.tmp, s_c[j], static_cast<int>(x), static_cast<int>(y));
            s_all[idx] = sum;
        }
        __syncthreads();
    }

    const int ntiles = tiles_for<kRDThreads, kRDNV>(ohw);
    for (int k = k_begin; k < k_end; ++k) {
        const int map = b * a.K + k;
        const size_t off = static_cast<size_t>(map) * ohw;
        const float* adv = a.y_adv ? a.y_adv + off : nullptr;
        const float* fz = has_fused ? a.fused + off : nullptr;
        const Centre ck = s_c[k];

        // ---- pass A: per-map max of the un-normalised ground-false map ------------------------
        float M = 1.0f;
        if (normalise) {
            if (TASK == RD_BWD) {
                M = a.stats[3 * map + 2];
            } else {
                Stats<3> sa;
                stats_init(sa);
                for (int tile = 0; tile < ntiles; ++tile) {
                    float4 f[kRDNV];
                    if (has_fused) load_tile<kRDThreads, kRDNV, MODE>(fz, ohw, tile, t, 0.0f, f);
#pragma unroll
                    for (int j = 0; j < kRDNV; ++j) {
                        const int idx0 = tile * (kRDThreads * kRDNV * 4) + (j * kRDThreads + t) * 4;
                        if (idx0 >= ohw) continue;
                        uint32_t y0, x0;
                        a.wdiv.divmod(static_cast<uint32_t>(idx0), y0, x0);
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            if (idx0 + c >= ohw) continue;
                            int x = static_cast<int>(x0) + c, y = static_cast<int>(y0);
                            if (MODE == WALK_SCALAR)
                                while (x >= a.ow) {
                                    x -= a.ow;
                                    ++y;
                                }
                            const float gt = patch_at(s_tab, a.tmp, ck, x, y);
                            const float g = ground_false_value(a.variant, has_fused, k, a.K, x, y, idx0 + c, gt,
                                                               has_fused ? f4_get(f[j], c) : 0.0f, s_tab, a.tmp, s_c, s_all);
                            sa.mx = nanmax(sa.mx, g);
                        }
                    }
                }
                group_reduce<kRDThreads, 3, false, false, true>(sa, scratch);
                M = sa.mx;
            }
        }

        // ---- pass B: the statistics / the gradient / the materialised maps ---------------------
        Stats<3> st;
        stats_init(st);
        float coef = 0.f, lb = 0.f, invS = 0.f;
        if (TASK == RD_BWD) {
            const float w = a.weight ? a.weight[map] : 1.0f;
            float go, denom;
            if (a.grad_kind == HP_GRAD_SCALAR) {
                go = a.grad_out[0];
                denom = static_cast<float>(a.B) * static_cast<float>(a.K);
            } else {
                go = a.grad_out[b];
                denom = static_cast<float>(a.K);
            }
            coef = go * w / denom;
            lb = -a.stats[3 * map + 0] * kLog2e;
            invS = 1.0f / a.stats[3 * map + 1];
        }
        for (int tile = 0; tile < ntiles; ++tile) {
            float4 p[kRDNV], f[kRDNV];
            if (TASK != RD_MATERIALIZE) load_tile<kRDThreads, kRDNV, MODE>(adv, ohw, tile, t, -INFINITY, p);
            if (has_fused && want_gf) load_tile<kRDThreads, kRDNV, MODE>(fz, ohw, tile, t, 0.0f, f);
            if (TASK == RD_FWD) softmax_tile<kRDNV>(st.m, st.s, p);
#pragma unroll
            for (int j = 0; j < kRDNV; ++j) {
                const int idx0 = tile * (kRDThreads * kRDNV * 4) + (j * kRDThreads + t) * 4;
                if (idx0 >= ohw) continue;
                uint32_t y0, x0;
                a.wdiv.divmod(static_cast<uint32_t>(idx0), y0, x0);
                float r_gt[4], r_out[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    r_gt[c] = 0.f;
                    r_out[c] = 0.f;
                    if (idx0 + c >= ohw) continue;
                    int x = static_cast<int>(x0) + c, y = static_cast<int>(y0);
                    if (MODE == WALK_SCALAR)
                        while (x >= a.ow) {
                            x -= a.ow;
                            ++y;
                        }
                    const float gt = patch_at(s_tab, a.tmp, ck, x, y);
                    float tgt = gt;
                    if (want_gf) {
                        float g = ground_false_value(a.variant, has_fused, k, a.K, x, y, idx0 + c, gt,
                                                     has_fused ? f4_get(f[j], c) : 0.0f, s_tab, a.tmp, s_c, s_all);
                        if (normalise) g = __fdiv_rn(g, M);  // 0/0 stays NaN like the reference (regda_7.py:3624)
                        tgt = g;
                    }
                    if (TASK == RD_FWD) {
                        kl_elem(st.sum, f4_get(p[j], c), tgt + a.eps);
                    } else if (TASK == RD_BWD) {
                        r_out[c] = coef * (exp2f(fmaf(f4_get(p[j], c), kLog2e, lb)) - (tgt + a.eps) * invS);
                    } else {
                        r_gt[c] = gt;
                        r_out[c] = tgt;
                    }
                }
                if (TASK == RD_BWD || TASK == RD_MATERIALIZE) {
                    float* o1 = (TASK == RD_BWD ? a.grad_in : a.gf) + off + idx0;
                    float* o2 = (TASK == RD_MATERIALIZE) ? a.gt + off + idx0 : nullptr;
                    if (MODE == WALK_VEC) {
                        stg_stream4(reinterpret_cast<float4*>(o1), make_float4(r_out[0], r_out[1], r_out[2], r_out[3]));
                        if (o2) stg_stream4(reinterpret_cast<float4*>(o2), make_float4(r_gt[0], r_gt[1], r_gt[2], r_gt[3]));
                    } else {
#pragma unroll
                        for (int c = 0; c < 4; ++c)
                            if (idx0 + c < ohw) {
                                o1[c] = r_out[c];
                                if (o2) o2[c] = r_gt[c];
                            }
                    }
                }
            }
        }
        if (TASK == RD_FWD) {
            group_reduce<kRDThreads, 3, false, true, false>(st, scratch);
            if (t == 0) {
                const float w = a.weight ? a.weight[map] : 1.0f;
                float lse;
                const double L = kl_finish(st.m, st.s, st.sum[0], st.sum[1], st.sum[2], lse);
                a.per_map[map] = static_cast<float>(L * static_cast<double>(w));
                a.stats[3 * map + 0] = lse;
                a.stats[3 * map + 1] = st.sum[0];
                a.stats[3 * map + 2] = M;
            }
        }
    }

    if (TASK == RD_FWD) {
        if (a.mean == nullptr && a.per_sample == nullptr) return;
        if (last_block_arrives(&a.ws->counter, gridDim.x)) {
            const int n_maps = a.B * a.K;
            if (a.per_sample) per_sample_means(a.per_map, a.B, a.K, a.per_sample, t, kRDThreads);
            if (a.mean) {
                const double acc = strided_sum_f64(a.per_map, n_maps, t, kRDThreads);
                s_red[t] = acc;
                __syncthreads();
                for (int o = kRDThreads / 2; o > 0; o >>= 1) {
                    if (t < o) s_red[t] += s_red[t + o];
                    __syncthreads();
                }
                if (t == 0) *a.mean = static_cast<float>(s_red[0] / static_cast<double>(n_maps));
            }
            if (t == 0) a.ws->counter = 0;
        }
    }
}

template <int TASK>
static int launch_regdisp(RDArgs a, bool vec, cudaStream_t stream, const char* who) {
    const int ohw = a.oh * a.ow;
    const bool want_gf = (TASK == RD_MATERIALIZE) || (a.mode == HP_MODE_MAX);
    const bool needs_all = want_gf && (a.variant == HP_RD_BASE || a.variant == HP_RD_X6 || a.variant == HP_RD_RD4);
    const int ntab = 2 * a.tmp * a.tmp + 1;
    const size_t smem = sizeof(float) * (((ntab + 3) & ~3) + (needs_all ? ohw : 0));
    HP_REQUIRE(smem <= 200 * 1024, HP_ERR_SHAPE, "%s: %dx%d map does not fit the per-sample shared-memory sum", who, a.oh,
               a.ow);
    // enough blocks for ~2 resident waves on 148 SMs x 3 blocks, never more than one block per joint
    int splits = (888 + a.B - 1) / a.B;
    splits = splits < 1 ? 1 : (splits > a.K ? a.K : splits);
    a.splits = splits;
    a.wdiv = FastDiv(static_cast<uint32_t>(a.ow));
    const int grid = a.B * splits;
    cudaError_t e = cudaSuccess;
    if (vec) {
        if (smem > 48 * 1024)
            e = cudaFuncSetAttribute(regdisp_kernel<WALK_VEC, TASK>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     static_cast<int>(smem));
        if (e == cudaSuccess) regdisp_kernel<WALK_VEC, TASK><<<grid, kRDThreads, smem, stream>>>(a);
    } else {
        if (smem > 48 * 1024)
            e = cudaFuncSetAttribute(regdisp_kernel<WALK_SCALAR, TASK>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     static_cast<int>(smem));
        if (e == cudaSuccess) regdisp_kernel<WALK_SCALAR, TASK><<<grid, kRDThreads, smem, stream>>>(a);
    }
    if (e != cudaSuccess) return fail(static_cast<int>(e), "%s: %s", who, cudaGetErrorString(e));
    return launch_status(who);
}

// HP_RD_SHAPE=g forces the guarded generic kernel (A/B comparisons; the staged kernel is the default)
static bool generic_forced() {
    const char* e = getenv("HP_RD_SHAPE");
    return e != nullptr && e[0] == 'g';
}

static int check_rd(const char* who, int variant, int mode, int B, int K, int oh, int ow, int tmp) {
    HP_REQUIRE(variant >= HP_RD_BASE && variant <= HP_RD_RD4, HP_ERR_ARG, "%s: variant %d", who, variant);
    HP_REQUIRE(mode == HP_MODE_MIN || mode == HP_MODE_MAX, HP_ERR_ARG, "%s: mode %d", who, mode);
    HP_REQUIRE(B > 0 && K > 0 && K <= HP_MAX_K && oh > 0 && ow > 0 && static_cast<long long>(oh) * ow < (1ll << 28),
               HP_ERR_SHAPE, "%s: bad shape B=%d K=%d oh=%d ow=%d", who, B, K, oh, ow);
    HP_REQUIRE(tmp >= 0 && tmp <= 64, HP_ERR_ARG, "%s: tmp %d", who, tmp);
    return HP_OK;
}

}  // namespace hp

using namespace hp;

// comparison runs: HP_RD_MIN_FUSED=0 keeps the two-launch path (decode, then the staged loss kernel) for 'min'
static bool rd_min_fused_enabled() {
    static const bool on = []() {
        const char* e = std::getenv("HP_RD_MIN_FUSED");
        return !(e && e[0] == '0');
    }();
    return on;
}

extern "C" HP_API int hp_regdisp_fwd(const float* y, const float* y_adv, const float* fused, const float* weight,
                                     int variant, int mode, float epsilon, int B, int K, int H, int W, int oh, int ow,
                                     int shift, int tmp, const float* tab, float* per_map, float* per_sample,
                                     float* mean, float* stats, int32_t* centres, void* workspace,
                                     hp_stream_t stream) {
    if (int rc = check_rd("hp_regdisp_fwd", variant, mode, B, K, oh, ow, tmp)) return rc;
    HP_REQUIRE(y && y_adv && tab && per_map && stats && centres && workspace, HP_ERR_NULL, "hp_regdisp_fwd: null pointer");
    HP_REQUIRE(H > 0 && W > 0 && shift >= 0 && shift < 16 && ((H - 1) >> shift) < oh && ((W - 1) >> shift) < ow,
               HP_ERR_SHAPE, "hp_regdisp_fwd: decoded %dx%d >> %d does not fit %dx%d", H, W, shift, oh, ow);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    // 'min' at full resolution needs nothing but the map's own argmax: decode + loss in ONE kernel (hp_regdisp_min.cuh)
    if (mode == HP_MODE_MIN && H == oh && W == ow && shift == 0 && H * W == 4096 && W % 4 == 0 && aligned16(y) &&
        aligned16(y_adv) && (2 * tmp + 1) * (2 * tmp + 1) <= 32 * kTileMaxPatch && !generic_forced() && rd_min_fused_enabled()) {
        RDMinArgs m{};
        m.y = y; m.y_adv = y_adv; m.weight = weight; m.n_maps = B * K; m.B = B; m.K = K; m.H = H; m.W = W; m.HW = H * W;
        m.tmp = tmp; m.wdiv = FastDiv(static_cast<uint32_t>(W)); m.sdiv = FastDiv(static_cast<uint32_t>(2 * tmp + 1));
        m.tab = tab; m.eps = epsilon;
        m.eps_log_eps = epsilon > 0.0f ? static_cast<float>(static_cast<double>(epsilon) * std::log(static_cast<double>(epsilon))) : 0.0f;
        m.per_map = per_map; m.per_sample = per_sample; m.mean = mean; m.stats = stats; m.centres = centres;
        m.ws = static_cast<Workspace*>(workspace);
        const int rc = launch_regdisp_min(m, s, "hp_regdisp_fwd");
        if (rc != 1) return rc;
    }
    // x1 / x5 recipes (constant target outside the own patch) on 32x32 / 16x16 heads, both modes, no fused map: decode +
    // loss in ONE kernel as well (hp_regdisp_sparse.cuh)
    const float* fused_used = (mode == HP_MODE_MAX && variant == HP_RD_X5) ? fused : nullptr;  // ('min' / x1 ignore y_adv2)
    if ((variant == HP_RD_X1 || variant == HP_RD_X5) && (fused_used == nullptr || aligned16(fused_used)) && H * W == 4096 && W % 4 == 0 &&
        (oh * ow == 1024 || oh * ow == 256) && ow % 4 == 0 && aligned16(y) && aligned16(y_adv) &&
        (2 * tmp + 1) * (2 * tmp + 1) <= 32 * kTileMaxPatch && (2 * tmp + 1) * (2 * tmp + 1) < oh * ow && !generic_forced() &&
        rd_min_fused_enabled()) {
        RDSparseArgs sp{};
        RDMinArgs& m = sp.m;
        m.y = y; m.y_adv = y_adv; m.weight = weight; m.n_maps = B * K; m.B = B; m.K = K; m.H = H; m.W = W; m.HW = H * W;
        m.tmp = tmp; m.wdiv = FastDiv(static_cast<uint32_t>(W)); m.sdiv = FastDiv(static_cast<uint32_t>(2 * tmp + 1));
        m.tab = tab; m.eps = epsilon;
        m.per_map = per_map; m.per_sample = per_sample; m.mean = mean; m.stats = stats; m.centres = centres;
        m.ws = static_cast<Workspace*>(workspace);
        sp.oh = oh; sp.ow = ow; sp.shift = shift; sp.want_gf = (mode == HP_MODE_MAX) ? 1 : 0; sp.bg = sp.want_gf ? 1.0f : 0.0f;
        sp.fused = fused_used;
        sp.ubg = sp.bg + epsilon;
        sp.ubg_log_ubg = sp.ubg > 0.0f ? sp.ubg * std::log(sp.ubg) : 0.0f;
        const int rc = launch_regdisp_sparse(sp, s, "hp_regdisp_fwd");
        if (rc != 1) return rc;
    }
    if (int rc = launch_decode(y, B * K, H, W, nullptr, nullptr, nullptr, centres, shift, s)) return rc;
    RDArgs a{};
    a.y_adv = y_adv; a.fused = fused; a.weight = weight; a.variant = variant; a.mode = mode; a.eps = epsilon;
    a.B = B; a.K = K; a.oh = oh; a.ow = ow; a.tmp = tmp; a.tab = tab; a.centres = centres;
    a.per_map = per_map; a.per_sample = per_sample; a.mean = mean; a.stats = stats;
    a.ws = static_cast<Workspace*>(workspace);
    const bool vec = (ow % 4 == 0) && aligned16(y_adv) && (!fused || aligned16(fused));
    if (vec && !generic_forced()) {
        int rc = launch_regdisp_dense<RD_FWD>(a, s, "hp_regdisp_fwd");  // x6 / rd4 'max' on 4096-pixel maps
        if (rc != 1) return rc;
        rc = launch_regdisp_staged<RD_FWD>(a, s, "hp_regdisp_fwd");
        if (rc != 1) return rc;
    }
    return launch_regdisp<RD_FWD>(a, vec, s, "hp_regdisp_fwd");
}

/* profiling aid: the blocks of the dense 'max' disparity kernel stamp their timeline (globaltimer ns) into `buf`
 * (device memory, hp_debug_regdisp_trace_words() uint64 words).  NULL switches it off.  Read by profiles/trace_regdisp.py. */
extern "C" HP_API size_t hp_debug_regdisp_trace_words(void) {
    int sms = hp_device_sm_count();
    if (sms <= 0) sms = 148;
    return static_cast<size_t>(sms) * kRDDTraceBlockWords;
}
extern "C" HP_API int hp_debug_regdisp_trace(void* buf, size_t words) {
    g_rdd_trace = static_cast<unsigned long long*>(buf);
    g_rdd_trace_words = buf ? words : 0;
    return HP_OK;
}

extern "C" HP_API int hp_regdisp_bwd(const float* y_adv, const float* fused, const float* weight, int variant, int mode,
                                     float epsilon, int B, int K, int oh, int ow, int tmp, const float* tab,
                                     const int32_t* centres, const float* stats, const float* grad_out, int grad_kind,
                                     float* grad_in, hp_stream_t stream) {
    if (int rc = check_rd("hp_regdisp_bwd", variant, mode, B, K, oh, ow, tmp)) return rc;
    HP_REQUIRE(y_adv && tab && centres && stats && grad_out && grad_in, HP_ERR_NULL, "hp_regdisp_bwd: null pointer");
    HP_REQUIRE(grad_kind == HP_GRAD_SCALAR || grad_kind == HP_GRAD_PER_SAMPLE, HP_ERR_ARG,
               "hp_regdisp_bwd: grad_kind %d", grad_kind);
    RDArgs a{};
    a.y_adv = y_adv; a.fused = fused; a.weight = weight; a.variant = variant; a.mode = mode; a.eps = epsilon;
    a.B = B; a.K = K; a.oh = oh; a.ow = ow; a.tmp = tmp; a.tab = tab; a.centres = centres; a.stats = const_cast<float*>(stats);
    a.grad_out = grad_out; a.grad_kind = grad_kind; a.grad_in = grad_in;
    const bool vec = (ow % 4 == 0) && aligned16(y_adv) && aligned16(grad_in) && (!fused || aligned16(fused));
    if (vec && !generic_forced()) {
        int rc = launch_regdisp_dense<RD_BWD>(a, static_cast<cudaStream_t>(stream), "hp_regdisp_bwd");  // x6 / rd4 'max', 4096-pixel maps
        if (rc != 1) return rc;
        rc = launch_regdisp_staged<RD_BWD>(a, static_cast<cudaStream_t>(stream), "hp_regdisp_bwd");
        if (rc != 1) return rc;
    }
    return launch_regdisp<RD_BWD>(a, vec, static_cast<cudaStream_t>(stream), "hp_regdisp_bwd");
}

// ---- x6 'max' with the fused map built IN the kernel from its two heads (train1.py:410-426; SURVEY.md 8d, configs[2] 'max' with
// in-kernel fusion: 37,888 B per map) -------------------------------------------------------------------------------------------
static int check_heads(const char* who, const float* f_lo, int hl, int wl, const float* f_mid, int hm, int wm, int variant, int mode,
                       int oh, int ow, int tmp, int K, const float* y_adv) {
    HP_REQUIRE(f_lo && f_mid, HP_ERR_NULL, "%s: null head", who);
    HP_REQUIRE(variant == HP_RD_X6 && mode == HP_MODE_MAX, HP_ERR_ARG,
               "%s: only RegressionDisparityx6 mode='max' reads a fused map (variant %d, mode %d)", who, variant, mode);
    HP_REQUIRE(oh == 64 && ow == 64 && hl == kRDDLoSide && wl == kRDDLoSide && hm == kRDDMidSide && wm == kRDDMidSide, HP_ERR_SHAPE,
               "%s: heads %dx%d / %dx%d -> %dx%d; the in-kernel fusion covers 16x16 / 32x32 -> 64x64 (materialise the map with "
               "hp_fuse_multiscale and call hp_regdisp_fwd otherwise)", who, hl, wl, hm, wm, oh, ow);
    HP_REQUIRE(K <= kRDDMaxK && tmp <= 6 && (2 * tmp + 1) * (2 * tmp + 1) <= 32 * kTileMaxPatch, HP_ERR_SHAPE,
               "%s: K=%d tmp=%d outside the dense kernel's range", who, K, tmp);
    HP_REQUIRE(aligned16(f_lo) && aligned16(f_mid) && aligned16(y_adv), HP_ERR_ARG, "%s: pointers must be 16-byte aligned", who);
    return HP_OK;
}

extern "C" HP_API int hp_regdisp_fwd_heads(const float* y, const float* y_adv, const float* f_lo, int hl, int wl, float a_lo,
                                           const float* f_mid, int hm, int wm, float a_mid, const float* weight, int variant,
                                           int mode, float epsilon, int B, int K, int H, int W, int oh, int ow, int tmp,
                                           const float* tab, float* per_map, float* per_sample, float* mean, float* stats,
                                           int32_t* centres, void* workspace, hp_stream_t stream) {
    if (int rc = check_rd("hp_regdisp_fwd_heads", variant, mode, B, K, oh, ow, tmp)) return rc;
    HP_REQUIRE(y && y_adv && tab && per_map && stats && centres && workspace, HP_ERR_NULL, "hp_regdisp_fwd_heads: null pointer");
    HP_REQUIRE(H == oh && W == ow, HP_ERR_SHAPE, "hp_regdisp_fwd_heads: y is %dx%d, the label grid %dx%d", H, W, oh, ow);
    if (int rc = check_heads("hp_regdisp_fwd_heads", f_lo, hl, wl, f_mid, hm, wm, variant, mode, oh, ow, tmp, K, y_adv)) return rc;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (int rc = launch_decode(y, B * K, H, W, nullptr, nullptr, nullptr, centres, 0, s)) return rc;
    RDArgs a{};
    a.y_adv = y_adv; a.f_lo = f_lo; a.f_mid = f_mid; a.a_lo = a_lo; a.a_mid = a_mid; a.weight = weight; a.variant = variant;
    a.mode = mode; a.eps = epsilon; a.B = B; a.K = K; a.oh = oh; a.ow = ow; a.tmp = tmp; a.tab = tab; a.centres = centres;
    a.per_map = per_map; a.per_sample = per_sample; a.mean = mean; a.stats = stats; a.ws = static_cast<Workspace*>(workspace);
    const int rc = launch_regdisp_dense<RD_FWD>(a, s, "hp_regdisp_fwd_heads");
    if (rc == 1) return fail(HP_ERR_ARG, "hp_regdisp_fwd_heads: the dense kernel is switched off (HP_RD_DENSE=0)");
    return rc;
}

extern "C" HP_API int hp_regdisp_bwd_heads(const float* y_adv, const float* f_lo, int hl, int wl, float a_lo, const float* f_mid,
                                           int hm, int wm, float a_mid, const float* weight, int variant, int mode, float epsilon,
                                           int B, int K, int oh, int ow, int tmp, const float* tab, const int32_t* centres,
                                           const float* stats, const float* grad_out, int grad_kind, float* grad_in,
                                           hp_stream_t stream) {
    if (int rc = check_rd("hp_regdisp_bwd_heads", variant, mode, B, K, oh, ow, tmp)) return rc;
    HP_REQUIRE(y_adv && tab && centres && stats && grad_out && grad_in, HP_ERR_NULL, "hp_regdisp_bwd_heads: null pointer");
    HP_REQUIRE(grad_kind == HP_GRAD_SCALAR || grad_kind == HP_GRAD_PER_SAMPLE, HP_ERR_ARG,
               "hp_regdisp_bwd_heads: grad_kind %d", grad_kind);
    if (int rc = check_heads("hp_regdisp_bwd_heads", f_lo, hl, wl, f_mid, hm, wm, variant, mode, oh, ow, tmp, K, y_adv)) return rc;
    HP_REQUIRE(aligned16(grad_in), HP_ERR_ARG, "hp_regdisp_bwd_heads: grad_in must be 16-byte aligned");
    RDArgs a{};
    a.y_adv = y_adv; a.f_lo = f_lo; a.f_mid = f_mid; a.a_lo = a_lo; a.a_mid = a_mid; a.weight = weight; a.variant = variant;
    a.mode = mode; a.eps = epsilon; a.B = B; a.K = K; a.oh = oh; a.ow = ow; a.tmp = tmp; a.tab = tab; a.centres = centres;
    a.stats = const_cast<float*>(stats); a.grad_out = grad_out; a.grad_kind = grad_kind; a.grad_in = grad_in;
    const int rc = launch_regdisp_dense<RD_BWD>(a, static_cast<cudaStream_t>(stream), "hp_regdisp_bwd_heads");
    if (rc == 1) return fail(HP_ERR_ARG, "hp_regdisp_bwd_heads: the dense kernel is switched off (HP_RD_DENSE=0)");
    return rc;
}

extern "C" HP_API int hp_regdisp_materialize(const float* fused, int variant, int B, int K, int oh, int ow, int tmp,
                                             const float* tab, const int32_t* centres, float* gt, float* gf,
                                             hp_stream_t stream) {
    if (int rc = check_rd("hp_regdisp_materialize", variant, HP_MODE_MAX, B, K, oh, ow, tmp)) return rc;
    HP_REQUIRE(tab && centres && gt && gf, HP_ERR_NULL, "hp_regdisp_materialize: null pointer");
    RDArgs a{};
    a.fused = fused; a.variant = variant; a.mode = HP_MODE_MAX; a.B = B; a.K = K; a.oh = oh; a.ow = ow; a.tmp = tmp;
    a.tab = tab; a.centres = centres; a.gt = gt; a.gf = gf;
    const bool vec = (ow % 4 == 0) && aligned16(gt) && aligned16(gf) && (!fused || aligned16(fused));
    return launch_regdisp<RD_MATERIALIZE>(a, vec, static_cast<cudaStream_t>(stream), "hp_regdisp_materialize");
}

extern "C" HP_API int hp_pseudo_label(const float* y, int B, int K, int H, int W, int kind, int oh, int ow, int shift,
                                      int tmp, const float* tab, float* gt, float* gf, int32_t* centres,
                                      hp_stream_t stream) {
    HP_REQUIRE(kind == HP_PLG_BASE || kind == HP_PLG_ONE_MINUS, HP_ERR_ARG, "hp_pseudo_label: kind %d", kind);
    const int variant = kind == HP_PLG_BASE ? HP_RD_BASE : HP_RD_X1;
    if (int rc = check_rd("hp_pseudo_label", variant, HP_MODE_MAX, B, K, oh, ow, tmp)) return rc;
    HP_REQUIRE(y && tab && gt && gf && centres, HP_ERR_NULL, "hp_pseudo_label: null pointer");
    HP_REQUIRE(H > 0 && W > 0 && shift >= 0 && shift < 16 && ((H - 1) >> shift) < oh && ((W - 1) >> shift) < ow,
               HP_ERR_SHAPE, "hp_pseudo_label: decoded %dx%d >> %d does not fit %dx%d", H, W, shift, oh, ow);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (int rc = launch_decode(y, B * K, H, W, nullptr, nullptr, nullptr, centres, shift, s)) return rc;
    return hp_regdisp_materialize(nullptr, variant, B, K, oh, ow, tmp, tab, centres, gt, gf, stream);
}
