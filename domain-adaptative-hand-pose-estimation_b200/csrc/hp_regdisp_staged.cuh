// hp_regdisp_staged.cuh - the production shape of the regression-disparity forward / backward kernels on sm_100a
// (a8-a11: RegressionDisparity{,x1,x5,x6} with the JointsKLLoss criterion; regda_4.py:129-143,
// regda_7.py:3250-3268, 3529-3561, 3609-3632).
//
// Shape (the pattern of the headline kernel, hp_pipeline_bulk.cuh):
//   * persistent blocks, each owning a CONTIGUOUS range of maps (balanced to one map over the grid);
//   * y_adv (and the fused map) of a map land in a shared-memory stage through the copy engine
//     (cp.async.bulk + mbarrier complete_tx); the ring of `kst` stages keeps ~100 KB per block requested
//     whatever the warps are doing; the threads only ever read shared memory, each one the same 4*NV pixels of
//     every map, which then live in registers for both passes (max, then sums against the true max);
//   * ONE block barrier per map: it publishes the per-warp maxima and frees the stage for the refill; the
//     per-map closure of map q runs on one thread after the barrier of map q+1 (double-buffered partials);
//   * the pseudo label is never a map: outside the joint's own (2 tmp + 1)^2 patch the target is a closed form
//     ('min': 0; x1 / x5 'max': 1) or a per-pixel function of the sample's summed Gaussians and the fused map
//     (x6 / base 'max', x5 with a fused map); the sum over the K joints is rebuilt once per sample in REGISTERS
//     (each thread for its own pixels, joints in ascending order -> deterministic, no barrier);
//     only float4s that touch the own patch take the exact per-pixel recipe;
//   * small per-sample inputs (centres, weights, backward coefficients) are fetched one sample ahead: a plain
//     load issued behind ~200 KB of queued bulk copies would wait microseconds.
// Algorithmic bytes per map: oh*ow*4 (y_adv) [+ oh*ow*4 fused] (+ H*W*4 for the decode launch that precedes
// the forward); backward adds oh*ow*4 written.  Roofline: HBM.
#pragma once
#include <cstdlib>

#include "hp_common.cuh"
#include "hp_internal.cuh"
#include "hp_tma.cuh"

namespace hp {

enum RDTask { RD_FWD = 0, RD_BWD = 1, RD_MATERIALIZE = 2 };

struct RDArgs {
    const float* y_adv;
    const float* fused;
    const float* weight;
    int variant, mode;
    float eps;
    int B, K, oh, ow, tmp;
    const float* tab;
    const int32_t* centres;
    int splits;
    FastDiv wdiv;
    // forward
    float* per_map;
    float* per_sample;
    float* mean;
    float* stats;  // [B*K,3] lse, S, M
    Workspace* ws;
    // backward
    const float* grad_out;
    int grad_kind;
    float* grad_in;
    // materialise
    float* gt;
    float* gf;
};

__device__ __forceinline__ float clip01(float x) { return (x != x) ? x : fminf(fmaxf(x, 0.0f), 1.0f); }

// un-normalised ground-false value of joint k at pixel (x, y).  `all` = clip01(sum over
// ALL joints of their Gaussians at this pixel) (base / x6 only), gt = joint k's own Gaussian, f = fused map.
__device__ __forceinline__ float ground_false_pixel(int variant, bool use_fused, int k, int K, int x, int y, float gt,
                                                    float f, float all, const float* s_tab, int tmp, const Centre* s_c) {
    float g;
    if (variant == HP_RD_BASE) {
        if (gt == 0.0f) return all;
        float sum = 0.0f;  // inside joint k's own patch: exclude it explicitly (no cancellation)
        for (int j = 0; j < K; ++j)
            if (j != k) sum += patch_at(s_tab, tmp, s_c[j], x, y);
        return clip01(sum);
    }
    if (variant == HP_RD_X6) g = clip01(__fsub_rn(all, __fmul_rn(gt, 10.0f)));
    else g = clip01(__fsub_rn(1.0f, __fmul_rn(gt, 10.0f)));
    if (use_fused) g = clip01(__fsub_rn(__fadd_rn(g, f), __fmul_rn(gt, 100.0f)));
    return g;
}

constexpr int kRDSMaxStages = 8;
struct RDSMeta {
    float a, b, c, d;  // forward: {weight}; backward: {coef, -lse*log2e, 1/S, M}
};
template <int NT>
struct RDSShared {
    uint64_t bars[kRDSMaxStages];
    Centre c[2][HP_MAX_K];
    RDSMeta meta[2][HP_MAX_K];
    float p1[2][NT / 32][2];  // per-warp {max p, max g} of the map in flight
    float p2[2][NT / 32][6];  // per-warp {sum exp, sum u, sum u p, sum u lg2 u, sum p over the closed-form pixels}
};

__device__ __forceinline__ float max4(float4 v) { return fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)); }
__device__ __forceinline__ float4 clip01_4(float4 v) { return make_float4(clip01(v.x), clip01(v.y), clip01(v.z), clip01(v.w)); }
__device__ __forceinline__ float4 add4(float4 a, float4 b) {
    return make_float4(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y), __fadd_rn(a.z, b.z), __fadd_rn(a.w, b.w));
}
__device__ __forceinline__ float warp_sum_f32(float x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
}
// u / M of the normalised recipes: exact division (bit-equal to the materialised maps), skipped when M == 1
__device__ __forceinline__ float rds_norm(float g, float M, bool divide) { return divide ? __fdiv_rn(g, M) : g; }

// NT threads per block, NV float4 per thread and map (NT * NV * 4 >= oh * ow), 512 / NT blocks per SM.
template <int NT, int NV, int TASK>
__global__ void __launch_bounds__(NT, 512 / NT) regdisp_staged_kernel(const RDArgs a, const int kst) {
    extern __shared__ __align__(128) unsigned char s_rds[];
    __shared__ RDSShared<NT> sh;
    __shared__ double s_red[TASK == RD_FWD ? NT : 1];
    constexpr int NW = NT / 32;
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int ohw = a.oh * a.ow, n4 = ohw >> 2, K = a.K, tmp = a.tmp;
    const bool want_gf = a.mode == HP_MODE_MAX;
    const bool use_fused = want_gf && a.fused != nullptr && a.variant != HP_RD_X1;
    const bool needs_all = want_gf && (a.variant == HP_RD_BASE || a.variant == HP_RD_X6);
    const bool normalise = want_gf && (a.variant == HP_RD_X5 || a.variant == HP_RD_X6);
    const bool dense = needs_all || use_fused;  // the target outside the own patch depends on the pixel
    const float bg = want_gf ? 1.0f : 0.0f;     // ... or is this constant (before normalisation and epsilon)
    const int nbuf = use_fused ? 2 : 1;
    const uint32_t stage_bytes = static_cast<uint32_t>(nbuf) * static_cast<uint32_t>(ohw) * 4u;
    const int ntab = 2 * tmp * tmp + 1;
    float* s_tab = reinterpret_cast<float*>(s_rds + static_cast<size_t>(kst) * stage_bytes);
    const uint32_t stage_u32 = smem_addr(s_rds), bar_u32 = smem_addr(sh.bars);

    // this block's maps: [m0, m1)
    const int n_maps = a.B * K;
    const int m0 = static_cast<int>((static_cast<long long>(blockIdx.x) * n_maps) / gridDim.x);
    const int m1 = static_cast<int>((static_cast<long long>(blockIdx.x + 1) * n_maps) / gridDim.x);
    const int q_total = m1 - m0;
    int sample = m0 / K, k = m0 - sample * K, spar = 0;

    // the pixels of this thread: float4 number t + j*NT of every map
    int gx[NV], gy[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        const int v = t + j * NT;
        uint32_t yy, xx;
        a.wdiv.divmod(static_cast<uint32_t>(v < n4 ? 4 * v : 0), yy, xx);
        gx[j] = static_cast<int>(xx);
        gy[j] = static_cast<int>(yy);
    }

    // per-sample inputs, one sample ahead (thread kk < K holds joint kk)
    auto load_meta = [&](int s, Centre& c, RDSMeta& mt) {
        if (t < K && s < a.B) {
            const int map = s * K + t;
            c.x = a.centres[2 * map + 0];
            c.y = a.centres[2 * map + 1];
            const float w = a.weight ? a.weight[map] : 1.0f;
            if (TASK == RD_FWD) {
                mt.a = w;
            } else {
                float go, denom;
                if (a.grad_kind == HP_GRAD_SCALAR) {
                    go = a.grad_out[0];
                    denom = static_cast<float>(a.B) * static_cast<float>(K);
                } else {
                    go = a.grad_out[s];
                    denom = static_cast<float>(K);
                }
                mt.a = go * w / denom;
                mt.b = -a.stats[3 * map + 0] * kLog2e;
                mt.c = 1.0f / a.stats[3 * map + 1];
                mt.d = a.stats[3 * map + 2];
            }
        }
    };
    auto request = [&](int q) {  // thread 0: arm stage q % kst and request map m0 + q
        const int s = q % kst;
        const size_t off = static_cast<size_t>(m0 + q) * ohw;
        mbar_arrive_expect_tx(bar_u32 + 8 * s, stage_bytes);
        const uint64_t pol = l2_evict_first_policy();
        bulk_load(stage_u32 + s * stage_bytes, a.y_adv + off, static_cast<uint32_t>(ohw) * 4u, bar_u32 + 8 * s, pol);
        if (use_fused)
            bulk_load(stage_u32 + s * stage_bytes + static_cast<uint32_t>(ohw) * 4u, a.fused + off,
                      static_cast<uint32_t>(ohw) * 4u, bar_u32 + 8 * s, pol);
    };

    // ---- prologue: small loads are ISSUED before the bulk copies are requested -------------------------------------
    Centre c_cur{0, 0}, c_nxt{0, 0};
    RDSMeta m_cur{0.f, 0.f, 0.f, 0.f}, m_nxt{0.f, 0.f, 0.f, 0.f};
    load_meta(sample, c_cur, m_cur);
    load_meta(sample + 1, c_nxt, m_nxt);
    const float tab0 = (t < ntab) ? a.tab[t] : 0.0f;
    if (t == 0) {
        for (int s = 0; s < kst; ++s) mbar_init(bar_u32 + 8 * s, 1);
        mbar_init_fence();
        for (int q = 0; q < kst && q < q_total; ++q) request(q);
    }
    if (t < ntab) s_tab[t] = tab0;
    for (int i = t + NT; i < ntab; i += NT) s_tab[i] = a.tab[i];
    if (t < K) {
        sh.c[0][t] = c_cur;
        sh.meta[0][t] = m_cur;
    }
    __syncthreads();

    float4 all[NV];  // clip01(sum over the sample's joints) at this thread's pixels
#pragma unroll
    for (int j = 0; j < NV; ++j) all[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    bool new_sample = true;
    // the map whose closure is pending (forward)
    int prev_map = -1;
    float prev_w = 0.f, prev_M = 1.f, prev_Mp = 0.f;
    Centre prev_c{0, 0};

    auto closure = [&](int pp) {  // ONE thread: per-map loss and the statistics the backward pass needs
        float Sexp = 0.f, Su = 0.f, Sup = 0.f, Sulg = 0.f, Spbg = 0.f;
        for (int w = 0; w < NW; ++w) {
            Sexp += sh.p2[pp][w][0];
            Su += sh.p2[pp][w][1];
            Sup += sh.p2[pp][w][2];
            Sulg += sh.p2[pp][w][3];
            Spbg += sh.p2[pp][w][4];
        }
        float Sulogu = Sulg * kLn2;
        if (!dense) {  // the pixels outside the float4s that touch the patch, in closed form
            const int r0 = max(prev_c.y - tmp, 0), r1 = min(prev_c.y + tmp, a.oh - 1);
            const int c0 = max(prev_c.x - tmp, 0) >> 2, c1 = min(prev_c.x + tmp, a.ow - 1) >> 2;
            const float nb = static_cast<float>(ohw - 4 * (r1 - r0 + 1) * (c1 - c0 + 1));
            const float ubg = rds_norm(bg, prev_M, normalise && prev_M != 1.0f) + a.eps;
            Su = fmaf(nb, ubg, Su);
            Sup = fmaf(ubg, Spbg, Sup);
            if (ubg != 0.0f) Sulogu = fmaf(nb, ubg * logf(ubg), Sulogu);
        }
        // L = (sum u ln u - sum u p)/S - ln S + lse,  lse = Mp + ln(sum exp)   (loss.py:145-158)
        const float lse = prev_Mp + logf(Sexp);
        const double L = static_cast<double>((Sulogu - Sup) / Su) + static_cast<double>(prev_Mp) +
                         static_cast<double>(logf(Sexp / Su));
        a.per_map[prev_map] = static_cast<float>(L * static_cast<double>(prev_w));
        a.stats[3 * prev_map + 0] = lse;
        a.stats[3 * prev_map + 1] = Su;
        a.stats[3 * prev_map + 2] = prev_M;
    };

    for (int q = 0; q < q_total; ++q) {
        const int map = m0 + q, par = q & 1, s = q % kst;
        const uint32_t phase = static_cast<uint32_t>(q / kst) & 1u;
        const Centre* s_c = sh.c[spar];
        if (new_sample) {
            new_sample = false;
            if (needs_all) {
#pragma unroll
                for (int j = 0; j < NV; ++j) {
                    float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (t + j * NT < n4) {
                        for (int jj = 0; jj < K; ++jj) {
                            const Centre cj = s_c[jj];
                            if (static_cast<unsigned>(gy[j] - cj.y + tmp) <= 2u * static_cast<unsigned>(tmp) &&
                                gx[j] + 3 >= cj.x - tmp && gx[j] <= cj.x + tmp)
                                sum = add4(sum, patch_at4(s_tab, tmp, cj, gx[j], gy[j]));
                        }
                    }
                    all[j] = clip01_4(sum);
                }
            }
        }
        if (k == K - 1 && t < K) {  // hand the next sample's inputs over (visible after this map's barrier)
            sh.c[spar ^ 1][t] = c_nxt;
            sh.meta[spar ^ 1][t] = m_nxt;
        }
        const Centre ck = s_c[k];
        const RDSMeta meta = sh.meta[spar][k];
        const float4* st_adv = reinterpret_cast<const float4*>(s_rds + static_cast<size_t>(s) * stage_bytes);
        const float4* st_fz = st_adv + n4;

        mbar_wait(bar_u32 + 8 * s, phase);
        // ---- pass A: the map into registers, the target's un-normalised values, the two maxima ----------------------
        float4 p[NV], g[NV];
        unsigned hit = 0;
        float lmp = -INFINITY, lmg = -INFINITY;
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            const int v = t + j * NT;
            p[j] = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
            g[j] = make_float4(bg, bg, bg, bg);
            if (v < n4) {
                p[j] = st_adv[v];
                float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
                if (use_fused) f = st_fz[v];
                if (dense) {
                    g[j] = needs_all ? all[j] : make_float4(1.f, 1.f, 1.f, 1.f);
                    if (use_fused) g[j] = clip01_4(add4(g[j], f));
                }
                const bool touches = static_cast<unsigned>(gy[j] - ck.y + tmp) <= 2u * static_cast<unsigned>(tmp) &&
                                     gx[j] + 3 >= ck.x - tmp && gx[j] <= ck.x + tmp;
                if (touches) {  // exact per-pixel recipe
                    hit |= 1u << j;
                    const float4 gt = patch_at4(s_tab, tmp, ck, gx[j], gy[j]);
                    if (want_gf) {
                        g[j].x = ground_false_pixel(a.variant, use_fused, k, K, gx[j] + 0, gy[j], gt.x, f.x, all[j].x, s_tab, tmp, s_c);
                        g[j].y = ground_false_pixel(a.variant, use_fused, k, K, gx[j] + 1, gy[j], gt.y, f.y, all[j].y, s_tab, tmp, s_c);
                        g[j].z = ground_false_pixel(a.variant, use_fused, k, K, gx[j] + 2, gy[j], gt.z, f.z, all[j].z, s_tab, tmp, s_c);
                        g[j].w = ground_false_pixel(a.variant, use_fused, k, K, gx[j] + 3, gy[j], gt.w, f.w, all[j].w, s_tab, tmp, s_c);
                    } else {
                        g[j] = gt;
                    }
                }
                lmp = fmaxf(lmp, max4(p[j]));
                lmg = fmaxf(lmg, max4(g[j]));
            }
        }
        if (TASK == RD_FWD) {
            const float wmp = warp_max_f32(lmp);
            const float wmg = normalise ? warp_max_f32(lmg) : 1.0f;
            if (lane == 0) {
                sh.p1[par][warp][0] = wmp;
                sh.p1[par][warp][1] = wmg;
            }
        }
        __syncthreads();  // every thread holds its pixels: the stage is free; the warp maxima are visible
        if (t == 0 && q + kst < q_total) request(q + kst);

        if (TASK == RD_FWD) {
            if (prev_map >= 0 && t == 32 * ((q - 1) % NW)) closure(par ^ 1);
            float Mp = -INFINITY, M = -INFINITY;
#pragma unroll
            for (int w = 0; w < NW; ++w) {
                Mp = fmaxf(Mp, sh.p1[par][w][0]);
                M = fmaxf(M, sh.p1[par][w][1]);
            }
            if (!normalise) M = 1.0f;
            const bool divide = normalise && M != 1.0f;
            // ---- pass B: the sums, against the true maximum -----------------------------------------------------------
            const float ms = (Mp == -INFINITY) ? 0.0f : Mp;
            const float mb = -ms * kLog2e;
            float sexp = 0.f, su = 0.f, sup = 0.f, sulg = 0.f, spbg = 0.f;
#pragma unroll
            for (int j = 0; j < NV; ++j) {
                if (t + j * NT < n4) {
                    sexp += ex2_approx(fmaf(p[j].x, kLog2e, mb)) + ex2_approx(fmaf(p[j].y, kLog2e, mb));
                    sexp += ex2_approx(fmaf(p[j].z, kLog2e, mb)) + ex2_approx(fmaf(p[j].w, kLog2e, mb));
                    if (dense || ((hit >> j) & 1u)) {
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            const float u = rds_norm(f4_get(g[j], c), M, divide) + a.eps;
                            const float pv = f4_get(p[j], c);
                            su += u;
                            sup = fmaf(u, pv, sup);
                            if (u != 0.0f) sulg = fmaf(u, lg2_approx(u), sulg);  // xlogy: 0 at u == 0, NaN for u < 0
                        }
                    } else {
                        spbg += (p[j].x + p[j].y) + (p[j].z + p[j].w);
                    }
                }
            }
            sexp = warp_sum_f32(sexp);
            su = warp_sum_f32(su);
            sup = warp_sum_f32(sup);
            sulg = warp_sum_f32(sulg);
            spbg = warp_sum_f32(spbg);
            if (lane == 0) {
                float* o = sh.p2[par][warp];
                o[0] = sexp;
                o[1] = su;
                o[2] = sup;
                o[3] = sulg;
                o[4] = spbg;
            }
            prev_map = map;
            prev_w = meta.a;
            prev_M = M;
            prev_Mp = Mp;
            prev_c = ck;
        } else {
            // ---- backward: d/dp = coef * (softmax(p) - u / S)    ----------------------------
            const float coef = meta.a, lb = meta.b, invS = meta.c, M = meta.d;
            const bool divide = normalise && M != 1.0f;
            const float ubg = rds_norm(bg, M, divide) + a.eps;
            float4* out = reinterpret_cast<float4*>(a.grad_in + static_cast<size_t>(map) * ohw);
#pragma unroll
            for (int j = 0; j < NV; ++j) {
                const int v = t + j * NT;
                if (v < n4) {
                    float r[4];
                    const bool exact = dense || ((hit >> j) & 1u);
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const float u = exact ? rds_norm(f4_get(g[j], c), M, divide) + a.eps : ubg;
                        r[c] = coef * (ex2_approx(fmaf(f4_get(p[j], c), kLog2e, lb)) - u * invS);
                    }
                    stg_stream4(out + v, make_float4(r[0], r[1], r[2], r[3]));
                }
            }
        }
        if (++k == K) {
            k = 0;
            ++sample;
            spar ^= 1;
            new_sample = true;
            load_meta(sample + 1, c_nxt, m_nxt);
        }
    }

    if (TASK == RD_FWD) {
        __syncthreads();
        if (prev_map >= 0 && t == 0) closure((q_total - 1) & 1);
        if (a.mean == nullptr && a.per_sample == nullptr) return;
        if (last_block_arrives(&a.ws->counter, gridDim.x)) {
            const volatile float* pmv = a.per_map;
            if (a.per_sample) {
                for (int s = t; s < a.B; s += NT) {
                    double acc = 0.0;
                    for (int kk = 0; kk < K; ++kk) acc += static_cast<double>(pmv[s * K + kk]);
                    a.per_sample[s] = static_cast<float>(acc / static_cast<double>(K));
                }
            }
            if (a.mean) {
                double acc = 0.0;
                for (int i = t; i < n_maps; i += NT) acc += static_cast<double>(pmv[i]);
                s_red[t] = acc;
                __syncthreads();
                for (int o = NT / 2; o > 0; o >>= 1) {
                    if (t < o) s_red[t] += s_red[t + o];
                    __syncthreads();
                }
                if (t == 0) *a.mean = static_cast<float>(s_red[0] / static_cast<double>(n_maps));
            }
            if (t == 0) a.ws->counter = 0;
        }
    }
}

// host side: pick the block shape, the stage count and the grid; returns 1 when the shape is not covered
// (the caller then takes the guarded generic kernel), 0 when launched, < 0 / > 0 on errors.
template <int NT, int NV, int TASK>
static int launch_rds_shape(const RDArgs& a, int nbuf, int sms, cudaStream_t stream, const char* who) {
    constexpr int BPS = 512 / NT;
    const int ohw = a.oh * a.ow;
    const size_t stage = static_cast<size_t>(nbuf) * ohw * 4;
    const size_t tab = ((table_bytes(a.tmp) + 15) / 16) * 16;
    const size_t budget = (227 * 1024) / BPS - 1024 - sizeof(RDSShared<NT>) - (TASK == RD_FWD ? NT * 8 : 8) - 256;
    if (budget < tab + 2 * stage) return 1;
    int kst = static_cast<int>((budget - tab) / stage);
    if (kst > kRDSMaxStages) kst = kRDSMaxStages;
    const size_t smem = static_cast<size_t>(kst) * stage + tab;
    const int n_maps = a.B * a.K;
    int grid = sms * BPS;
    if (const char* e = getenv("HP_RD_GRID")) {  // tests: few blocks -> long map ranges spanning many samples
        const int g = atoi(e);
        if (g > 0) grid = g;
    }
    if (grid > n_maps) grid = n_maps;
    static bool attr_done_dev[64] = {};  // per instantiation and device
    int dev = 0;
    cudaGetDevice(&dev);
    bool& attr_done = attr_done_dev[dev & 63];
    if (!attr_done) {
        const cudaError_t e = cudaFuncSetAttribute(regdisp_staged_kernel<NT, NV, TASK>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                   static_cast<int>((227 * 1024) / BPS - 1024 - 4096));
        if (e != cudaSuccess) return fail(static_cast<int>(e), "%s: %s", who, cudaGetErrorString(e));
        attr_done = true;
    }
    regdisp_staged_kernel<NT, NV, TASK><<<grid, NT, smem, stream>>>(a, kst);
    return launch_status(who);
}

template <int TASK>
static int launch_regdisp_staged(RDArgs a, cudaStream_t stream, const char* who) {
    const int ohw = a.oh * a.ow, n4 = ohw / 4;
    if (n4 > 1024) return 1;
    static int sms = 0;
    if (sms == 0) {
        sms = hp_device_sm_count();
        if (sms <= 0) sms = 148;
    }
    a.wdiv = FastDiv(static_cast<uint32_t>(a.ow));
    const bool use_fused = a.mode == HP_MODE_MAX && a.fused != nullptr && a.variant != HP_RD_X1;
    const int nbuf = use_fused ? 2 : 1;
    if (n4 <= 64) return launch_rds_shape<64, 1, TASK>(a, nbuf, sms, stream, who);
    if (n4 <= 256) return launch_rds_shape<128, 2, TASK>(a, nbuf, sms, stream, who);
    return launch_rds_shape<256, 4, TASK>(a, nbuf, sms, stream, who);
}

}  // namespace hp
