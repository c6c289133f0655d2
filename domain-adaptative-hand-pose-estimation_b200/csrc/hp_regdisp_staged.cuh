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
    // dense kernel only: the fused map as its two heads, fused = a_lo up64(f_lo) + a_mid up64(f_mid) (16x16 / 32x32 per map)
    const float* f_lo;
    const float* f_mid;
    float a_lo, a_mid;
    const float* weight;
    int variant, mode;
    float eps;
    int B, K, oh, ow, tmp;
    const float* tab;
    const int32_t* centres;
    int splits;
    FastDiv wdiv;
    FastDiv sdiv;  // by the patch side (dense kernel)
    unsigned producer_sleep_ns;  // staged kernels: back-off of the producer warp's poll on an empty-barrier
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
    // profiling (hp_debug_regdisp_trace): per block kRDDTraceBlockWords stamps of the dense kernel, or null
    unsigned long long* trace;
};

// clamp to [0, 1], NaN propagates (torch.clamp); two FMNMX.NAN
__device__ __forceinline__ float clip01(float x) {
    float y;
    asm("max.NaN.f32 %0, %1, 0f00000000;\n\tmin.NaN.f32 %0, %0, 0f3F800000;" : "=f"(y) : "f"(x));
    return y;
}

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
    if (variant == HP_RD_X6 || variant == HP_RD_RD4) g = clip01(__fsub_rn(all, __fmul_rn(gt, 10.0f)));
    else g = clip01(__fsub_rn(1.0f, __fmul_rn(gt, 10.0f)));
    if (use_fused) g = clip01(__fsub_rn(__fadd_rn(g, f), __fmul_rn(gt, 100.0f)));
    return g;
}

constexpr int kRDSMaxStages = 8;
constexpr int kRDSBatch = 16;  // maps per closure batch
constexpr int kRDSSlots = 4;   // samples whose per-joint inputs are resident in shared memory
struct RDSMeta {
    float a, b, c, d;  // forward: {weight}; backward: {coef, -lse*log2e, 1/S, M}
};
template <int NT>
struct RDSShared {
    uint64_t full[kRDSMaxStages];   // copy landed (armed with the byte count by the producer)
    uint64_t empty[kRDSMaxStages];  // every consumer warp has read the stage out
    Centre c[kRDSSlots][HP_MAX_K];
    RDSMeta meta[kRDSSlots][HP_MAX_K];
    float mx[2][NT / 32];                       // normalised recipes: per-warp max of the un-normalised target
    float cl[2][kRDSBatch][NT / 32][8];         // per map and warp {max p, sum exp, sum u, sum u p, sum u lg2 u, sum p (bg)}
    float4 clm[2][kRDSBatch];                   // per map {centre x, centre y (bits), weight, M}
    unsigned long long acc[kFxAccWords];        // block sums of the per-map losses (fx_acc_add)
};



__device__ __forceinline__ float max4(float4 v) { return fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)); }
__device__ __forceinline__ float4 clip01_4(float4 v) { return make_float4(clip01(v.x), clip01(v.y), clip01(v.z), clip01(v.w)); }
__device__ __forceinline__ float4 add4(float4 a, float4 b) {
    return make_float4(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y), __fadd_rn(a.z, b.z), __fadd_rn(a.w, b.w));
}
// Sum 8 per-lane values over the warp in 9 shuffles (transpose-reduce): every lane of the group of four with
// index i = lane >> 2 returns the total of v[4*bit4(lane) + 2*bit3(lane) + bit2(lane)], i.e. v[rds_sum_index(lane)].
__device__ __forceinline__ float warp_sum8_scattered(const float (&v)[8], int lane) {
    const bool h16 = (lane & 16) != 0, h8 = (lane & 8) != 0, h4 = (lane & 4) != 0;
    float w[4], x[2];
#pragma unroll
    for (int i = 0; i < 4; ++i) w[i] = (h16 ? v[i + 4] : v[i]) + __shfl_xor_sync(0xffffffffu, h16 ? v[i] : v[i + 4], 16);
#pragma unroll
    for (int i = 0; i < 2; ++i) x[i] = (h8 ? w[i + 2] : w[i]) + __shfl_xor_sync(0xffffffffu, h8 ? w[i] : w[i + 2], 8);
    float y = (h4 ? x[1] : x[0]) + __shfl_xor_sync(0xffffffffu, h4 ? x[0] : x[1], 4);
    y += __shfl_xor_sync(0xffffffffu, y, 2);
    y += __shfl_xor_sync(0xffffffffu, y, 1);
    return y;
}
__device__ __forceinline__ int rds_sum_index(int lane) { return ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1); }
// u / M of the normalised recipes: exact division (bit-equal to the materialised maps), skipped when M == 1
__device__ __forceinline__ float rds_norm(float g, float M, bool divide) { return divide ? __fdiv_rn(g, M) : g; }

// NT consumer threads + one producer warp per block, NV float4 per consumer thread and map (NT * NV * 4 == oh * ow),
// 512 / NT blocks per SM.  DENSE: the target outside the own patch depends on the pixel (x6 / base 'max', fused map);
// otherwise it is a constant there ('min': 0; x1 / x5 'max': 1) and the sums over those pixels are closed forms.
// blocks per SM: the sparse recipes are light in registers and take a third block of 256 threads (the kernels are
// issue-latency bound: more resident warps, not more bytes in flight, is what they lack)
template <int NT, bool DENSE>
constexpr int rds_blocks_per_sm() { return (NT == 256 && !DENSE) ? 3 : 512 / NT; }

template <int NT, int NV, int TASK, bool DENSE, bool FUSED>
__global__ void __launch_bounds__(NT + 32, rds_blocks_per_sm<NT, DENSE>()) regdisp_staged_kernel(const RDArgs a, const int kst) {
    extern __shared__ __align__(128) unsigned char s_rds[];
    __shared__ RDSShared<NT> sh;
    constexpr int NW = NT / 32;
    constexpr int n4 = NT * NV, ohw = 4 * n4;
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int K = a.K, tmp = a.tmp;
    const bool want_gf = a.mode == HP_MODE_MAX;
    const bool needs_all = DENSE && (a.variant == HP_RD_BASE || a.variant == HP_RD_X6 || a.variant == HP_RD_RD4);
    const bool normalise = DENSE && (a.variant == HP_RD_X5 || a.variant == HP_RD_X6);  // sparse recipes: M == 1 (host)
    const float bg = want_gf ? 1.0f : 0.0f;  // the constant target outside the patch of the sparse recipes
    constexpr uint32_t stage_bytes = (FUSED ? 2u : 1u) * static_cast<uint32_t>(ohw) * 4u;
    const int ntab = 2 * tmp * tmp + 1;
    float* s_tab = reinterpret_cast<float*>(s_rds + static_cast<size_t>(kst) * stage_bytes);
    const uint32_t stage_u32 = smem_addr(s_rds), full_u32 = smem_addr(sh.full), empty_u32 = smem_addr(sh.empty);

    // this block's maps: [m0, m1)
    const int n_maps = a.B * K;
    const int m0 = static_cast<int>((static_cast<long long>(blockIdx.x) * n_maps) / gridDim.x);
    const int m1 = static_cast<int>((static_cast<long long>(blockIdx.x + 1) * n_maps) / gridDim.x);
    const int q_total = m1 - m0;
    const int sample0 = m0 / K, k0 = m0 - sample0 * K;

    if (warp == NW) {
        // =================================== producer warp ==========================================================
        // lane l holds the per-sample inputs of joints l and l + 32 of the NEXT sample to be published; the loads
        // are issued a whole sample ahead (a plain load behind ~100 KB of queued bulk copies waits microseconds)
        Centre c[2];
        RDSMeta mt[2];
        auto load_meta = [&](int s) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int kk = lane + 32 * h;
                c[h] = Centre{0, 0};
                mt[h] = RDSMeta{0.f, 0.f, 0.f, 0.f};
                if (kk < K && s < a.B) {
                    const int map = s * K + kk;
                    c[h].x = a.centres[2 * map + 0];
                    c[h].y = a.centres[2 * map + 1];
                    const float w = a.weight ? a.weight[map] : 1.0f;
                    if (TASK == RD_FWD) {
                        mt[h].a = w;
                    } else {
                        float go, denom;
                        if (a.grad_kind == HP_GRAD_SCALAR) {
                            go = a.grad_out[0];
                            denom = static_cast<float>(a.B) * static_cast<float>(K);
                        } else {
                            go = a.grad_out[s];
                            denom = static_cast<float>(K);
                        }
                        mt[h].a = go * w / denom;
                        mt[h].b = -a.stats[3 * map + 0] * kLog2e;
                        mt[h].c = 1.0f / a.stats[3 * map + 1];
                        mt[h].d = a.stats[3 * map + 2];
                    }
                }
            }
        };
        auto publish = [&](int slot) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int kk = lane + 32 * h;
                if (kk < K) {
                    sh.c[slot][kk] = c[h];
                    sh.meta[slot][kk] = mt[h];
                }
            }
        };
        const uint64_t pol = l2_evict_first_policy();
        auto request = [&](int q) {  // lane 0: arm stage q % kst and request map m0 + q
            const int s = q % kst;
            const size_t off = static_cast<size_t>(m0 + q) * ohw;
            mbar_arrive_expect_tx(full_u32 + 8 * s, stage_bytes);
            bulk_load(stage_u32 + s * stage_bytes, a.y_adv + off, static_cast<uint32_t>(ohw) * 4u, full_u32 + 8 * s, pol);
            if (FUSED)
                bulk_load(stage_u32 + s * stage_bytes + static_cast<uint32_t>(ohw) * 4u, a.fused + off,
                          static_cast<uint32_t>(ohw) * 4u, full_u32 + 8 * s, pol);
        };
        // prologue: small loads first, then the first sample's maps, then its inputs
        load_meta(sample0);
        int q = min(min(kst, K - k0), q_total);
        if (lane == 0) {
            for (int s = 0; s < kst; ++s) {
                mbar_init(full_u32 + 8 * s, 1);
                mbar_init(empty_u32 + 8 * s, NW);
            }
            mbar_init_fence();
            for (int i = 0; i < q; ++i) request(i);
        }
        publish(0);
        load_meta(sample0 + 1);
        __syncthreads();
        int kq = k0 + q, smp = sample0;
        if (kq == K) {
            kq = 0;
            ++smp;
        }
        for (; q < q_total; ++q) {
            if (kq == 0) {  // first map of a sample: its inputs become visible with the arm of the map's barrier
                publish((smp - sample0) & (kRDSSlots - 1));
                __syncwarp();
                load_meta(smp + 1);
            }
            if (lane == 0) {
                if (q >= kst) mbar_wait_backoff(empty_u32 + 8 * (q % kst), static_cast<uint32_t>(q / kst - 1) & 1u, a.producer_sleep_ns);
                request(q);
            }
            __syncwarp();
            if (++kq == K) {
                kq = 0;
                ++smp;
            }
        }
    } else {
        // =================================== consumer warps =========================================================
        for (int i = t; i < ntab; i += NT) s_tab[i] = a.tab[i];
        if (t < kFxAccWords) sh.acc[t] = 0ull;
        __syncthreads();  // barriers initialised, table and the first sample's inputs in shared memory
        // the pixels of this thread: float4 number t + j*NT of every map
        int gx[NV], gy[NV];
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            uint32_t yy, xx;
            a.wdiv.divmod(static_cast<uint32_t>(4 * (t + j * NT)), yy, xx);
            gx[j] = static_cast<int>(xx);
            gy[j] = static_cast<int>(yy);
        }
        float4 all[DENSE ? NV : 1];  // clip01(sum over the sample's joints) at this thread's pixels (1 when not needed)
#pragma unroll
        for (int j = 0; j < (DENSE ? NV : 1); ++j) all[j] = make_float4(1.f, 1.f, 1.f, 1.f);
        bool new_sample = true;
        int k = k0, smp = sample0, s = 0;
        uint32_t phase = 0;
        const unsigned rspan = 2u * static_cast<unsigned>(tmp), cspan = rspan + 3u;
        const float2 l2 = make_float2(kLog2e, kLog2e);

        auto close_batch = [&](int buf, int count, int first_map) {  // thread (warp i % NW, lane i / NW) closes slot i
            named_barrier<1, NT>();
            const int i = lane * NW + warp;
            if (i < count) {
                const float4 cm = sh.clm[buf][i];
                const int cx = __float_as_int(cm.x), cy = __float_as_int(cm.y);
                const float w = cm.z, M = cm.w;
                float Mp = -INFINITY;
                for (int ww = 0; ww < NW; ++ww) Mp = fmaxf(Mp, sh.cl[buf][i][ww][5]);
                float Sexp = 0.f, Su = 0.f, Sup = 0.f, Sulg = 0.f, Spbg = 0.f;
                for (int ww = 0; ww < NW; ++ww) {
                    const float* r = sh.cl[buf][i][ww];
                    Sexp = fmaf(r[0], exp_diff(r[5], Mp), Sexp);
                    Su += r[1];
                    Sup += r[2];
                    Sulg += r[3];
                    Spbg += r[4];
                }
                float Sulogu = Sulg * kLn2;
                if (!DENSE) {  // the pixels outside the float4s that touch the patch, in closed form
                    const int r0 = max(cy - tmp, 0), r1 = min(cy + tmp, a.oh - 1);
                    const int c0 = max(cx - tmp, 0) >> 2, c1 = min(cx + tmp, a.ow - 1) >> 2;
                    const float nb = static_cast<float>(ohw - 4 * (r1 - r0 + 1) * (c1 - c0 + 1));
                    const float ubg = bg + a.eps;
                    Su = fmaf(nb, ubg, Su);
                    Sup = fmaf(ubg, Spbg, Sup);
                    if (ubg != 0.0f) Sulogu = fmaf(nb, ubg * logf(ubg), Sulogu);
                }
                // L = (sum u ln u - sum u p)/S - ln S + lse,  lse = Mp + ln(sum exp)   (loss.py:145-158)
                const float lse = Mp + logf(Sexp);
                const double L = static_cast<double>((Sulogu - Sup) / Su) + static_cast<double>(Mp) +
                                 static_cast<double>(logf(Sexp / Su));
                const int map = first_map + i;
                const float Lw = static_cast<float>(L * static_cast<double>(w));
                a.per_map[map] = Lw;
                if (a.mean) fx_acc_add(sh.acc, Lw);
                a.stats[3 * map + 0] = lse;
                a.stats[3 * map + 1] = Su;
                a.stats[3 * map + 2] = M;
            }
        };
        // exact per-pixel target (before normalisation and epsilon) of a float4 that touches the own patch
        auto exact4 = [&](int j, const Centre ck, const Centre* s_c, float4 f, float4 al) {
            const float4 gt = patch_at4(s_tab, tmp, ck, gx[j], gy[j]);
            if (!want_gf) return gt;
            float4 r;
            r.x = ground_false_pixel(a.variant, FUSED, k, K, gx[j] + 0, gy[j], gt.x, f.x, al.x, s_tab, tmp, s_c);
            r.y = ground_false_pixel(a.variant, FUSED, k, K, gx[j] + 1, gy[j], gt.y, f.y, al.y, s_tab, tmp, s_c);
            r.z = ground_false_pixel(a.variant, FUSED, k, K, gx[j] + 2, gy[j], gt.z, f.z, al.z, s_tab, tmp, s_c);
            r.w = ground_false_pixel(a.variant, FUSED, k, K, gx[j] + 3, gy[j], gt.w, f.w, al.w, s_tab, tmp, s_c);
            return r;
        };

        for (int q = 0; q < q_total; ++q) {
            const int map = m0 + q, slot = (smp - sample0) & (kRDSSlots - 1);
            const Centre* s_c = sh.c[slot];
            const float4* st_adv = reinterpret_cast<const float4*>(s_rds + static_cast<size_t>(s) * stage_bytes);
            mbar_wait(full_u32 + 8 * s, phase);
            if (new_sample) {
                new_sample = false;
                if (needs_all) {
#pragma unroll
                    for (int j = 0; j < (DENSE ? NV : 1); ++j) {
                        float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
                        for (int jj = 0; jj < K; ++jj) {
                            const Centre cj = s_c[jj];
                            if (static_cast<unsigned>(gy[j] - cj.y + tmp) <= rspan &&
                                static_cast<unsigned>(gx[j] - cj.x + tmp + 3) <= cspan)
                                sum = add4(sum, patch_at4(s_tab, tmp, cj, gx[j], gy[j]));
                        }
                        all[j] = clip01_4(sum);
                    }
                }
            }
            const Centre ck = s_c[k];
            const RDSMeta meta = sh.meta[slot][k];
            const int r0 = ck.y - tmp, c0 = ck.x - tmp - 3;
            // ---- pass A: the map into registers, the target's un-normalised values (dense recipes), the maxima ------
            float4 p[NV], g[DENSE ? NV : 1];
            unsigned hit = 0;
            float lmp = -INFINITY, lmg = -INFINITY;
#pragma unroll
            for (int j = 0; j < NV; ++j) {
                p[j] = st_adv[t + j * NT];
                const bool touches = static_cast<unsigned>(gy[j] - r0) <= rspan && static_cast<unsigned>(gx[j] - c0) <= cspan;
                if (touches) hit |= 1u << j;
                if (DENSE) {
                    float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (FUSED) f = st_adv[n4 + t + j * NT];
                    g[j] = all[j];
                    if (FUSED) g[j] = clip01_4(add4(g[j], f));
                    if (touches) g[j] = exact4(j, ck, s_c, f, all[j]);
                    lmg = fmaxf(lmg, max4(g[j]));
                }
                lmp = fmaxf(lmp, max4(p[j]));
            }
            // every lane holds its pixels: this warp is done with the stage
            __syncwarp();
            if (lane == 0) mbar_arrive(empty_u32 + 8 * s);
            if (++s == kst) {
                s = 0;
                phase ^= 1u;
            }

            float M = 1.0f;
            if (TASK == RD_FWD) {
                if (DENSE && normalise) {  // the one cross-warp dependency of a map: the normaliser
                    const float wmg = warp_max_f32(lmg);
                    if (lane == 0) sh.mx[q & 1][warp] = wmg;
                    named_barrier<1, NT>();
                    M = -INFINITY;
#pragma unroll
                    for (int ww = 0; ww < NW; ++ww) M = fmaxf(M, sh.mx[q & 1][ww]);
                    if (M != 1.0f) {
#pragma unroll
                        for (int j = 0; j < (DENSE ? NV : 1); ++j)
                            g[j] = make_float4(__fdiv_rn(g[j].x, M), __fdiv_rn(g[j].y, M), __fdiv_rn(g[j].z, M), __fdiv_rn(g[j].w, M));
                    }
                }
                // ---- pass B: the sums; the softmax sum against the WARP's maximum (re-based at the closure) ----------
                const float wmp = warp_max_f32(lmp);
                const float ms = (wmp == -INFINITY) ? 0.0f : wmp;
                const float2 mb2 = make_float2(-ms * kLog2e, -ms * kLog2e);
                float2 sexp2 = make_float2(0.f, 0.f), spbg2 = make_float2(0.f, 0.f);
                float su = 0.f, sup = 0.f, sulg = 0.f;
#pragma unroll
                for (int j = 0; j < NV; ++j) {
                    const float2 lo = make_float2(p[j].x, p[j].y), hi = make_float2(p[j].z, p[j].w);
                    const float2 a0 = __ffma2_rn(lo, l2, mb2), a1 = __ffma2_rn(hi, l2, mb2);
                    sexp2 = __fadd2_rn(sexp2, __fadd2_rn(make_float2(ex2_approx(a0.x), ex2_approx(a0.y)),
                                                         make_float2(ex2_approx(a1.x), ex2_approx(a1.y))));
                    if (DENSE || ((hit >> j) & 1u)) {
                        const float4 gv = DENSE ? g[j] : exact4(j, ck, s_c, make_float4(0.f, 0.f, 0.f, 0.f), make_float4(1.f, 1.f, 1.f, 1.f));
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            const float u = f4_get(gv, c) + a.eps;
                            su += u;
                            sup = fmaf(u, f4_get(p[j], c), sup);
                            // xlogy: the term is 0 at u == 0 (lg2 of the clamp is finite), NaN propagates through u
                            sulg = fmaf(u, lg2_approx(fmaxf(u, 1.17549435e-38f)), sulg);
                        }
                    } else {
                        spbg2 = __fadd2_rn(spbg2, __fadd2_rn(lo, hi));
                    }
                }
                const float v8[8] = {sexp2.x + sexp2.y, su, sup, sulg, spbg2.x + spbg2.y, 0.f, 0.f, 0.f};
                const float tot = warp_sum8_scattered(v8, lane);
                const int buf = (q / kRDSBatch) & 1, i = q % kRDSBatch;
                if ((lane & 3) == 0) {
                    const int idx = rds_sum_index(lane);
                    if (idx < 6) sh.cl[buf][i][warp][idx] = (idx == 5) ? wmp : tot;
                }
                if (t == 0) sh.clm[buf][i] = make_float4(__int_as_float(ck.x), __int_as_float(ck.y), meta.a, M);
                if (i == kRDSBatch - 1 || q == q_total - 1) close_batch(buf, i + 1, map - i);
            } else {
                // ---- backward: d/dp = coef * (softmax(p) - u / S)    --------------------------
                const float coef = meta.a, lb = meta.b, invS = meta.c;
                M = meta.d;
                const bool divide = DENSE && normalise && M != 1.0f;
                const float qbg = (bg + a.eps) * invS;
                float4* out = reinterpret_cast<float4*>(a.grad_in + static_cast<size_t>(map) * ohw);
#pragma unroll
                for (int j = 0; j < NV; ++j) {
                    float r[4];
                    float4 gv = make_float4(bg, bg, bg, bg);
                    if (DENSE) gv = g[j];
                    else if ((hit >> j) & 1u) gv = exact4(j, ck, s_c, make_float4(0.f, 0.f, 0.f, 0.f), make_float4(1.f, 1.f, 1.f, 1.f));
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        float qv = qbg;
                        if (DENSE || ((hit >> j) & 1u)) qv = (rds_norm(f4_get(gv, c), M, divide) + a.eps) * invS;
                        r[c] = coef * (ex2_approx(fmaf(f4_get(p[j], c), kLog2e, lb)) - qv);
                    }
                    stg_stream4(out + t + j * NT, make_float4(r[0], r[1], r[2], r[3]));
                }
            }
            if (++k == K) {
                k = 0;
                ++smp;
                new_sample = true;
            }
        }
    }

    if (TASK == RD_FWD) {
        if (a.mean == nullptr && a.per_sample == nullptr) return;
        __syncthreads();  // the block's closures are done (shared accumulators final)
        if (a.mean && t < kFxAccWords && sh.acc[t] != 0ull) atomicAdd(&a.ws->acc[t], sh.acc[t]);
        if (last_block_arrives(&a.ws->counter, gridDim.x)) {
            if (a.per_sample) per_sample_means(a.per_map, a.B, K, a.per_sample, t, NT + 32);
            if (a.mean && t == 0) {
                *a.mean = fx_mean_from_workspace(a.ws->acc, n_maps);
            }
            if (t == 0) a.ws->counter = 0;
        }
    }
}

// host side: pick the block shape, the stage count and the grid; returns 1 when the shape is not covered
// (the caller then takes the guarded generic kernel), 0 when launched, < 0 / > 0 on errors.
template <int NT, int NV, int TASK, bool DENSE, bool FUSED>
static int launch_rds_shape(const RDArgs& a, int sms, cudaStream_t stream, const char* who) {
    constexpr int BPS = rds_blocks_per_sm<NT, DENSE>();
    const size_t stage = static_cast<size_t>(FUSED ? 2 : 1) * NT * NV * 16;
    const size_t tab = ((table_bytes(a.tmp) + 15) / 16) * 16;
    const size_t budget = (227 * 1024) / BPS - 1024 - sizeof(RDSShared<NT>) - 256;
    if (budget < tab + 2 * stage) return 1;
    int kst = static_cast<int>((budget - tab) / stage);
    if (kst > kRDSMaxStages) kst = kRDSMaxStages;
    if (kst > 2 * a.K) kst = 2 * a.K;  // the producer never runs more than two samples ahead (kRDSSlots)
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
        const cudaError_t e = cudaFuncSetAttribute(regdisp_staged_kernel<NT, NV, TASK, DENSE, FUSED>,
                                                   cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(budget));
        if (e != cudaSuccess) return fail(static_cast<int>(e), "%s: %s", who, cudaGetErrorString(e));
        attr_done = true;
    }
    regdisp_staged_kernel<NT, NV, TASK, DENSE, FUSED><<<grid, NT + 32, smem, stream>>>(a, kst);
    return launch_status(who);
}

template <int NT, int NV, int TASK>
static int launch_rds_recipe(const RDArgs& a, bool dense, bool fused, int sms, cudaStream_t stream, const char* who) {
    if (fused) return launch_rds_shape<NT, NV, TASK, true, true>(a, sms, stream, who);
    if (dense) return launch_rds_shape<NT, NV, TASK, true, false>(a, sms, stream, who);
    return launch_rds_shape<NT, NV, TASK, false, false>(a, sms, stream, who);
}

template <int TASK>
static int launch_regdisp_staged(RDArgs a, cudaStream_t stream, const char* who) {
    const int ohw = a.oh * a.ow;
    if (ohw != 256 && ohw != 1024 && ohw != 4096) return 1;  // 16^2, 32^2, 64^2 maps (any oh x ow with ow % 4 == 0)
    static int sms = 0;
    if (sms == 0) {
        sms = hp_device_sm_count();
        if (sms <= 0) sms = 148;
    }
    a.wdiv = FastDiv(static_cast<uint32_t>(a.ow));
    a.producer_sleep_ns = 200;
    if (const char* e = getenv("HP_RD_SLEEP")) a.producer_sleep_ns = static_cast<unsigned>(atoi(e));  // comparison runs
    const bool want_gf = a.mode == HP_MODE_MAX;
    const bool fused = want_gf && a.fused != nullptr && a.variant != HP_RD_X1 && a.variant != HP_RD_RD4;
    const bool dense = fused || (want_gf && (a.variant == HP_RD_BASE || a.variant == HP_RD_X6 || a.variant == HP_RD_RD4));
    if (!dense) {
        // sparse recipes take M == 1 for granted: some float4 of the map must lie outside every possible patch
        const int side = 2 * a.tmp + 1;
        const long long touched = static_cast<long long>(side < a.oh ? side : a.oh) * ((side + 2) / 4 + 1);
        if (touched >= ohw / 4) return 1;
    }
    if (ohw == 256) return launch_rds_recipe<64, 1, TASK>(a, dense, fused, sms, stream, who);
    if (ohw == 1024) return launch_rds_recipe<128, 2, TASK>(a, dense, fused, sms, stream, who);
    return launch_rds_recipe<256, 4, TASK>(a, dense, fused, sms, stream, who);
}

}  // namespace hp
