// hp_regdisp_dense.cuh - RegressionDisparityx6 / RegressionDisparity4 / RegressionDisparity (base) with mode='max' on
// 4096-pixel maps, forward and backward
// (configs[2] of BASELINE.json: 512 x 21 x 64 x 64; regda_7.py:3609-3632, regda_4.py RD4), with or without the fused
// map y_adv2 (train1.py:419-421).  The ground-false label of joint k,
//     lp   = clip(sum_j G_j)                      per SAMPLE (G_j = Gaussian at joint j's decoded centre)
//     g_k  = clip(lp - 10 G_k)                    [+ fused map f_k:  g_k = clip((g_k + f_k) - 100 G_k)]
//     gf_k = g_k / max(g_k),  u = gf_k + eps,     KL(y_adv_k, u)  (loss.py:145-158)
// depends on the pixel everywhere ("dense"), which made the register-slice kernel (hp_regdisp_staged.cuh) issue-bound:
// 5,000 warp-instructions per map, 0.47 / 0.60 of the HBM roofline.  This kernel splits the label into what is shared
// and what is not:
//   * outside joint k's own (2 tmp + 1)^2 patch G_k = 0, so g_k = lp (no fused map) or clip(lp + f_k): `lp` is built
//     ONCE per sample, in shared memory, by a builder warp running a sample ahead of the consumers (scatter of the K
//     patches, joints ascending -> deterministic; two slots, full/empty mbarriers);
//   * the own patch (<= 192 pixels) is a per-lane correction: the exact recipe is evaluated there from `lp`, the fused map
//     and the tabulated Gaussian, and the generic contribution of those pixels is taken back out;
//   * without a fused map the max of g_k (x6; rd4 does not normalise at all) is 1 whenever some other joint's centre lies
//     outside k's patch (lp = 1 there, G_k = 0); then sum(u) and sum(u lg u) are PER-SAMPLE constants plus the patch
//     correction, and the only per-pixel work of a map is the softmax sum and sum(lp * p).  Maps where every other centre
//     falls inside the own patch take an exact per-pixel path (rare; tested);
//   * with a fused map a first pass writes g = clip(lp + f) over f in the slot and finds the max, the second pass
//     accumulates the sums of u = g / M + eps.
// Shape: one block per SM; NG groups of G consumer warps, a group owns NS slots of one map (16 KB) each, filled by the copy
// engine (cp.async.bulk -> mbarrier complete_tx) - the group's load sequence (y_adv of its maps, or fused map / y_adv
// alternating) goes round the slots, so the next unit is in flight while the current map is computed.  The G warps of
// a group split the map's 32 iterations (and the patch slots) and meet at named barriers: one per map without a fused
// map (each warp runs the softmax against its own maximum, the leader re-bases), three with one (patch pixels marked,
// maxima exchanged, sums exchanged).  The leader role rotates over the group's warps.  No fused map: 6 groups x 2 warps x
// 2 slots; fused: 4 groups x 4 warps x 3 slots; + 2 x 16 KB of `lp`; contiguous map ranges per block (balanced to one
// map; ranges cross samples).  History (profiles/r2_regdisp_dense_history.md): one warp per map ran at 0.3 instructions
// per cycle and warp - 2.8 us (no fused map) / 3.7 us (fused) per map whatever the copies did.
// Algorithmic bytes per map: 16,384 (y_adv) [+ 16,384 fused] here + 16,384 (y) in the decode launch that precedes it.
// Roofline: HBM.
#pragma once
#include "hp_pipeline_parts.cuh"
#include "hp_regdisp_staged.cuh"
#include "hp_bilinear_block.cuh"

namespace hp {

constexpr int kRDDMaxK = 32;     // joints per sample (lane j of the builder holds joint j)
constexpr int kRDDMaxTab = 80;   // 2 tmp^2 + 1 <= 73 table entries (tmp <= 6)
constexpr int kRDDPixels = 4096;
// profiling: per block 8 header words {entry, exit, first lp built, n_samples, q_total, ...} (globaltimer ns), then per
// consumer warp (<= 16) and map (first 16) eight stamps {wait begins, data landed, lp seen, patch done, passes done,
// sums reduced + next units requested, map closed, -}, then per sample
// (first 8) two builder stamps {build begins, lp published}
constexpr int kRDDTraceMaps = 16;
constexpr int kRDDTraceBlockWords = 8 + 16 * kRDDTraceMaps * 8 + 8 * 2;
__device__ __forceinline__ unsigned long long rdd_now() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// LZ: 0 = the fused map (if any) arrives pre-fused; 4 / 5 = built in the kernel from its heads (LAZY) with 4 / 5 groups.
// Measured (B200, 512 x 21 maps, forward incl. the decode launch): 4 groups x 2 head slots 112.1 us; 5 groups x 1 head slot
// (672 threads: 80 registers, 152 B of spills) 114.2 us - a map takes a group 4.4 instead of 3.6 us - so 4 it is.
constexpr int kRDDLazyGroups = 4;
template <bool FUSED, int GW, int LZ = 0>
struct RDDShape {
    static constexpr int NG = LZ ? LZ : (FUSED ? 4 : 6);     // groups = maps being computed at once
    static constexpr int G = GW;                             // warps per group
    static constexpr int W = NG * G;                         // consumer warps
    static_assert(NG * GW <= 20, "RDDShared::xch / xmx hold 20 consumer warps");
    static constexpr int NS = FUSED ? 3 : 2;                 // slots per group
    static constexpr int UPM = FUSED ? 2 : 1;                // units per map
    static constexpr int IT = 32 / G;                        // iterations per warp and pass
    static constexpr int NSL = (kTileMaxPatch + G - 1) / G;  // patch slots per lane of a consumer warp
    static constexpr size_t kSmem = static_cast<size_t>(NG) * NS * kRDDPixels * 4 + 2 * kRDDPixels * 4;
    // LAZY (the fused map built from its heads): a group owns two 16 KB slots for y_adv and 5 KB slots for the heads, so that
    // the group's next map is in flight while the current one is computed (with a pre-fused map the 3 x 16 KB rotation leaves
    // the next map's y_adv unrequested until the current map closes: 1.0 us of every map's 3.1 us is a wait for it, 0.13 us
    // here).  LZ = 4: two head slots.  LZ = 5 (kept for comparison): the forward's heads are dead after pass 1, ONE head slot
    // re-requested behind the barrier that follows it, and the shared memory saved holds a fifth group.
    static constexpr int kHeadSlots = (LZ == 5) ? 1 : 2;
    static constexpr size_t kLazyGroupBytes = 2 * kRDDPixels * 4 + kHeadSlots * (16 * 16 + 32 * 32) * 4;
    static constexpr size_t kLazySmem = static_cast<size_t>(NG) * kLazyGroupBytes + 2 * kRDDPixels * 4;
};

struct RDDShared {
    uint64_t full[16];        // slot s of group g landed: full[g * NS + s]  (LAZY: full[4 g + {0, 1}] y_adv, full[4 g + {2, 3}] heads)
    uint64_t lp_full[2];      // lp slot built (count 1: the builder)
    uint64_t lp_empty[2];     // lp slot released (count W: every consumer warp, once per sample)
    Centre c[2][kRDDMaxK];
    float w[2][kRDDMaxK];
    float cu[2], culg[2];     // per sample: sum (lp + eps), sum (lp + eps) lg2 (lp + eps)
    float tab[kRDDMaxTab];
    union {
        float xch[2][20][8];          // forward: [map parity][consumer warp]: the warp's partial sums of a map
        RDSMeta meta[2][kRDDMaxK];    // backward: per sample slot and joint {coef, -lse log2 e, 1 / S, M}
    };
    float xmx[2][20];         // fused: max g of the warp's part
    unsigned long long acc[kFxAccWords];
};

// per-warp exact accumulation of the per-map losses (lane 0's registers; fx_acc_add's arithmetic, one set of shared
// atomics per warp at the end instead of per map)
struct FxReg {
    long long hi, lo;
    int n_nan, n_pinf, n_ninf;
};
__device__ __forceinline__ void fx_reg_add(FxReg& r, float v) {
    if (fabsf(v) < kFxAccLimit) {
        const double d = static_cast<double>(v) * 1099511627776.0;  // exact
        const long long hi = __double2ll_rn(d);
        r.hi += hi;
        r.lo += __double2ll_rn((d - static_cast<double>(hi)) * 8388608.0);
    } else if (v != v) {
        r.n_nan += 1;
    } else if (v > 0.0f) {
        r.n_pinf += 1;
    } else {
        r.n_ninf += 1;
    }
}
__device__ __forceinline__ void fx_reg_flush(const FxReg& r, unsigned long long* acc) {
    if (r.hi != 0) atomicAdd(&acc[0], static_cast<unsigned long long>(r.hi));
    if (r.lo != 0) atomicAdd(&acc[4], static_cast<unsigned long long>(r.lo));
    if (r.n_nan != 0) atomicAdd(&acc[1], static_cast<unsigned long long>(r.n_nan));
    if (r.n_pinf != 0) atomicAdd(&acc[2], static_cast<unsigned long long>(r.n_pinf));
    if (r.n_ninf != 0) atomicAdd(&acc[3], static_cast<unsigned long long>(r.n_ninf));
}
__device__ __forceinline__ void group_barrier(int id, int n_threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n_threads) : "memory");
}

__device__ __forceinline__ float ulg2(float u) { return u * lg2_approx(fmaxf(u, 1.17549435e-38f)); }  // xlogy / ln 2
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// the per-sample constants sum (lp + eps), sum (lp + eps) lg2 (lp + eps) are summed in ONE order whichever path built
// the sample (builder warp alone, or all warps for a block's first sample): per lane over the iterations ascending, then
// the fixed shuffle tree - these are the per-(iteration, lane) terms
__device__ __forceinline__ float rdd_term_u(float4 v, float eps) { return ((v.x + eps) + (v.y + eps)) + ((v.z + eps) + (v.w + eps)); }
__device__ __forceinline__ float rdd_term_ulg(float4 v, float eps) {
    return (ulg2(v.x + eps) + ulg2(v.y + eps)) + (ulg2(v.z + eps) + ulg2(v.w + eps));
}

// un-normalised ground-false value of joint k at an own-patch pixel, no fused map: x6 / rd4 clip(lp - 10 t); base
// clip(sum_{j != k} G_j) with the own Gaussian left out explicitly (no cancellation; regda_4.py:83-84)
__device__ __forceinline__ float rdd_patch_value(int variant, float lpv, float t, int k, int K, int x, int y, const float* s_tab, int tmp,
                                                 const Centre* s_c) {
    if (variant == HP_RD_BASE) {
        float sum = 0.0f;
        for (int j = 0; j < K; ++j)
            if (j != k) sum += patch_at(s_tab, tmp, s_c[j], x, y);
        return clip01(sum);
    }
    return clip01(__fsub_rn(lpv, __fmul_rn(t, 10.0f)));
}

// per-lane patch slots, lane-constant for the whole kernel: offset from the centre (dx = 1 << 20: unused) and value
// slot j of a warp that takes every `stride`-th slot row starting at `first` is patch pixel (first + j * stride) * 32 + lane
template <int N>
struct RDDSlots {
    int dx[N], dy[N];
    float t[N];
};
template <int N>
__device__ __forceinline__ void rdd_slots_init(RDDSlots<N>& s, const float* __restrict__ tab, int tmp, FastDiv sdiv, int lane, int first, int stride) {
    const int side = 2 * tmp + 1, n_patch = side * side;
#pragma unroll
    for (int k = 0; k < N; ++k) {
        const int i = (first + k * stride) * 32 + lane;
        s.dx[k] = 1 << 20;
        s.dy[k] = 0;
        s.t[k] = 0.0f;
        if (i < n_patch) {
            uint32_t qy, qx;
            sdiv.divmod(static_cast<uint32_t>(i), qy, qx);
            const int ry = static_cast<int>(qy), rx = static_cast<int>(qx);
            s.dx[k] = rx - tmp;
            s.dy[k] = ry - tmp;
            s.t[k] = tab[s.dx[k] * s.dx[k] + s.dy[k] * s.dy[k]];
        }
    }
}

// Exact per-pixel evaluation of a map without a fused map (every other centre lies inside the own patch, so the
// maximum of g is not known in closed form).  Returns the sums of u = g / M + eps (sum u, sum u (p - ref) log2 e with
// mb = -ref log2 e, sum u lg2 u) and M; two passes over lp.
__device__ __noinline__ void rdd_exact_map(const float4* __restrict__ P, const float4* __restrict__ LP, const float* s_tab,
                                           int tmp, Centre ck, FastDiv wdiv, float eps, float mb, int lane, float& Su,
                                           float& Sua, float& Sulg, float& M_out) {
    float mg = -INFINITY;
    for (int it = 0; it < 32; ++it) {
        const int i4 = it * 32 + lane;
        uint32_t yy, xx;
        wdiv.divmod(static_cast<uint32_t>(4 * i4), yy, xx);
        const float4 gt = patch_at4(s_tab, tmp, ck, static_cast<int>(xx), static_cast<int>(yy));
        const float4 l = LP[i4];
        mg = fmaxf(mg, fmaxf(fmaxf(clip01(__fsub_rn(l.x, __fmul_rn(gt.x, 10.0f))), clip01(__fsub_rn(l.y, __fmul_rn(gt.y, 10.0f)))),
                             fmaxf(clip01(__fsub_rn(l.z, __fmul_rn(gt.z, 10.0f))), clip01(__fsub_rn(l.w, __fmul_rn(gt.w, 10.0f))))));
    }
    const float M = warp_max_f32(mg);
    float su = 0.f, sup = 0.f, sulg = 0.f;
    for (int it = 0; it < 32; ++it) {
        const int i4 = it * 32 + lane;
        uint32_t yy, xx;
        wdiv.divmod(static_cast<uint32_t>(4 * i4), yy, xx);
        const float4 gt = patch_at4(s_tab, tmp, ck, static_cast<int>(xx), static_cast<int>(yy));
        const float4 l = LP[i4], p = P[i4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const float g = clip01(__fsub_rn(f4_get(l, c), __fmul_rn(f4_get(gt, c), 10.0f)));
            const float u = rds_norm(g, M, M != 1.0f) + eps;
            su += u;
            sup = fmaf(u, fmaf(f4_get(p, c), kLog2e, mb), sup);  // u * (p - ref) * log2 e
            sulg += ulg2(u);
        }
    }
    const float r = warp_sum3_scattered(su, sup, sulg, lane);
    Su = __shfl_sync(0xffffffffu, r, 0);
    Sua = __shfl_sync(0xffffffffu, r, 8);
    Sulg = __shfl_sync(0xffffffffu, r, 16);
    M_out = M;
}

// Softmax sum of iterations [it0, it1) of a map in shared memory against their TRUE maximum (two passes).  The passes
// of the kernel run against a sampled reference value instead (the maximum of the part's first iteration: the sum is
// the same number up to rounding whatever the reference, as long as nothing overflows) and come here when their sum is
// not a positive finite number.
__device__ __noinline__ void rdd_softmax_exact(const float4* __restrict__ P4, int it0, int it1, int lane, float& m_out, float& s_out) {
    float run = -INFINITY;
    for (int it = it0; it < it1; ++it) run = fmaxf(run, max4(P4[it * 32 + lane]));
    const float m = warp_max_f32(run);
    const float mb = -((m == -INFINITY) ? 0.0f : m) * kLog2e;
    float acc = 0.0f;
    for (int it = it0; it < it1; ++it) {
        const float4 v = P4[it * 32 + lane];
        acc += ex2_approx(fmaf(v.x, kLog2e, mb)) + ex2_approx(fmaf(v.y, kLog2e, mb));
        acc += ex2_approx(fmaf(v.z, kLog2e, mb)) + ex2_approx(fmaf(v.w, kLog2e, mb));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    m_out = m;
    s_out = acc;
}

// ---- the fused map built in the kernel (LAZY; hp_regdisp_fwd_heads / hp_regdisp_bwd_heads) ---------------------------------------
// train1.py:410-426 feeds x6 'max' with target5 = a_lo up64(y_adv3) + a_mid up64(y_adv2) (16x16 and 32x32 heads,
// nn.Upsample(mode='bilinear')).  Pre-fused, that map costs 16 KB written by the fusion kernel and 16 KB read here per
// map; LAZY launches receive the two heads themselves (1 KB + 4 KB per map, one bulk copy each into the head of the slot the
// fused map would have occupied) and every lane interpolates the fused values of its two 4x4 output blocks in registers -
// the static tap pattern of hp_bilinear_block.cuh, same arithmetic and operation order as the fusion kernels, so the values
// are bit-identical to hp_fuse_multiscale's.  Algorithmic bytes per map: 16,384 (y, decode) + 16,384 (y_adv) + 4,096 + 1,024
// = 37,888 (SURVEY.md 8d, configs[2] 'max' with in-kernel fusion) instead of 49,152 + the fusion kernel's 5,120 + 16,384.
constexpr int kRDDLoSide = 16, kRDDMidSide = 32;
constexpr int kRDDRow4 = 16;  // float4 per row of the 64 x 64 map the heads are fused into
constexpr int kRDDLoBytes = kRDDLoSide * kRDDLoSide * 4, kRDDMidBytes = kRDDMidSide * kRDDMidSide * 4;

// taps of output coordinate X at the exact scale S (align_corners=False: src = max(X / S + 0.5 / S - 0.5, 0), exact in fp32 for
// S in {2, 4}; i0 = floor(src), i1 = min(i0 + 1, in - 1), l1 = src - i0) - the scalar form of BlockAxis, branch-free so that the
// patch pixels' four bilinear gathers interleave
template <int S>
__device__ __forceinline__ void rdd_exact_tap(int X, int in, int& i0, int& i1, float& l1) {
    const float src = fmaxf(fmaf(static_cast<float>(X), 1.0f / S, 0.5f / S - 0.5f), 0.0f);
    i0 = __float2int_rz(src);
    l1 = src - static_cast<float>(i0);
    i1 = min(i0 + 1, in - 1);
}
template <int S>
__device__ __forceinline__ float rdd_bilinear_at(const float* __restrict__ src, int in, int x, int y) {
    int x0, x1, y0, y1;
    float lx, ly;
    rdd_exact_tap<S>(x, in, x0, x1, lx);
    rdd_exact_tap<S>(y, in, y0, y1, ly);
    const float lx0 = 1.0f - lx, ly0 = 1.0f - ly;
    const float t0 = __fmaf_rn(lx, src[y0 * in + x1], __fmul_rn(lx0, src[y0 * in + x0]));
    const float t1 = __fmaf_rn(lx, src[y1 * in + x1], __fmul_rn(lx0, src[y1 * in + x0]));
    return __fmaf_rn(ly, t1, __fmul_rn(ly0, t0));
}
// one pixel of the fused map from the staged heads (own-patch pixels)
__device__ __forceinline__ float rdd_fused_at(const float* __restrict__ lo, const float* __restrict__ mid, float a_lo, float a_mid, int x, int y) {
    return __fmaf_rn(a_mid, rdd_bilinear_at<2>(mid, kRDDMidSide, x, y), __fmul_rn(a_lo, rdd_bilinear_at<4>(lo, kRDDLoSide, x, y)));
}
// the lane's two 4x4 blocks (column block n, block rows m_begin, m_begin + 1): f[4 i + u] = row 4 (m_begin + i) + u, columns 4n..4n+3
__device__ __forceinline__ void rdd_fused_blocks(uint32_t lo_s, uint32_t mid_s, float a_lo, float a_mid, int n, int m_begin, float4 (&f)[8]) {
    const BlockAxis<4> lo_x = block_axis<4>(n);
    const BlockCols<4> lo_c = block_cols<4>(n, kRDDLoSide);
    const BlockAxis<2> mid_x = block_axis<2>(n);
    const BlockCols<2> mid_c = block_cols<2>(n, kRDDMidSide);
    BlockRows<4> RL;
    BlockRows<2> RM;
    block_rows_start<4>(RL, lo_s, kRDDLoSide, kRDDLoSide, m_begin, lo_c, lo_x);
    block_rows_start<2>(RM, mid_s, kRDDMidSide, kRDDMidSide, m_begin, mid_c, mid_x);
    const float2 al = make_float2(a_lo, a_lo), am = make_float2(a_mid, a_mid);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        float2 vl[4][2], vm[4][2];
        block_rows_blend<4>(RL, lo_s, kRDDLoSide, kRDDLoSide, m_begin + i, lo_c, lo_x, vl);
        block_rows_blend<2>(RM, mid_s, kRDDMidSide, kRDDMidSide, m_begin + i, mid_c, mid_x, vm);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const float2 r0 = __ffma2_rn(am, vm[u][0], __fmul2_rn(al, vl[u][0]));
            const float2 r1 = __ffma2_rn(am, vm[u][1], __fmul2_rn(al, vl[u][1]));
            f[4 * i + u] = make_float4(r0.x, r0.y, r1.x, r1.y);
        }
    }
}

// the sums of pass 2 with a fused map (u = g / M + eps), one float4 of the prediction and of g at a time
struct RDDFusedAcc {
    float2 s2, su2, sup2, suq2, sulg2, sulh2;
};
__device__ __forceinline__ void rdd_fused_acc(RDDFusedAcc& A, float4 v, float4 g, float2 l2e, float2 mb2, float2 im2, float2 e2) {
    const float2 lo = make_float2(v.x, v.y), hi = make_float2(v.z, v.w);
    const float2 a0 = __ffma2_rn(lo, l2e, mb2), a1 = __ffma2_rn(hi, l2e, mb2);
    A.s2 = __fadd2_rn(A.s2, __fadd2_rn(make_float2(ex2_approx(a0.x), ex2_approx(a0.y)), make_float2(ex2_approx(a1.x), ex2_approx(a1.y))));
    const float2 u0 = __ffma2_rn(make_float2(g.x, g.y), im2, e2), u1 = __ffma2_rn(make_float2(g.z, g.w), im2, e2);
    A.su2 = __fadd2_rn(A.su2, __fadd2_rn(u0, u1));
    A.sup2 = __ffma2_rn(u0, a0, A.sup2);
    A.suq2 = __ffma2_rn(u1, a1, A.suq2);
    const float2 q0 = make_float2(lg2_approx(fmaxf(u0.x, 1.17549435e-38f)), lg2_approx(fmaxf(u0.y, 1.17549435e-38f)));
    const float2 q1 = make_float2(lg2_approx(fmaxf(u1.x, 1.17549435e-38f)), lg2_approx(fmaxf(u1.y, 1.17549435e-38f)));
    A.sulg2 = __ffma2_rn(u0, q0, A.sulg2);
    A.sulh2 = __ffma2_rn(u1, q1, A.sulh2);
}

template <bool FUSED, int GW, int TASK, bool LAZY = false>
__global__ void __launch_bounds__(32 * (RDDShape<FUSED, GW, LAZY ? kRDDLazyGroups : 0>::W + 1), 1) regdisp_dense_kernel(const RDArgs a) {
    extern __shared__ __align__(128) unsigned char s_rdd[];
    __shared__ RDDShared sh;
    using S = RDDShape<FUSED, GW, LAZY ? kRDDLazyGroups : 0>;
    constexpr int kHeadSlots = S::kHeadSlots;
    constexpr int W = S::W, G = S::G, NG = S::NG, NS = S::NS, UPM = S::UPM, IT = S::IT, NSL = S::NSL;
    constexpr int kMapBytes = kRDDPixels * 4;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int K = a.K, tmp = a.tmp, ow = a.ow, oh = a.oh;
    const float eps = a.eps;
    constexpr size_t kGroupBytes = LAZY ? S::kLazyGroupBytes : static_cast<size_t>(NS) * kMapBytes;
    constexpr int kGroupBars = LAZY ? 2 + kHeadSlots : NS;  // LAZY: y_adv slots 0 / 1, then the head slot(s)
    constexpr int kHeadBytes = kRDDLoBytes + kRDDMidBytes;
    float* lp_base = reinterpret_cast<float*>(s_rdd + static_cast<size_t>(NG) * kGroupBytes);

    // this block's maps: [m0, m1), its samples: [sample0, sample0 + n_samples)
    const int n_maps = a.B * K;
    const int m0 = static_cast<int>((static_cast<long long>(blockIdx.x) * n_maps) / gridDim.x);
    const int m1 = static_cast<int>((static_cast<long long>(blockIdx.x + 1) * n_maps) / gridDim.x);
    const int q_total = m1 - m0;
    const int sample0 = m0 / K;
    const int n_samples = (q_total > 0) ? (m1 - 1) / K - sample0 + 1 : 0;
    unsigned long long* trace = a.trace ? a.trace + static_cast<size_t>(blockIdx.x) * kRDDTraceBlockWords : nullptr;
    if (trace && threadIdx.x == 0) {
        trace[0] = rdd_now();
        trace[3] = static_cast<unsigned long long>(n_samples);
        trace[4] = static_cast<unsigned long long>(q_total);
    }

    // the builder's first loads go out before any bulk copy is queued (a plain load behind ~200 KB of queued copies
    // per SM waits microseconds, and the first lp is on every consumer's critical path)
    Centre cn = Centre{0, 0};
    float wn = 1.0f;
    RDSMeta mn = RDSMeta{0.f, 0.f, 0.f, 1.f};  // backward only
    auto load_sample = [&](int s) {
        if (lane < K && s < a.B) {
            const int map = s * K + lane;
            cn.x = a.centres[2 * map + 0];
            cn.y = a.centres[2 * map + 1];
            wn = a.weight ? a.weight[map] : 1.0f;
            if (TASK == RD_BWD) {  // d/dp = coef (softmax(p) - u / S), coef = grad_out w / (B K) resp. / K   (appendix A6)
                const bool scalar = a.grad_kind == HP_GRAD_SCALAR;
                const float go = scalar ? a.grad_out[0] : a.grad_out[s];
                const float denom = scalar ? static_cast<float>(a.B) * static_cast<float>(K) : static_cast<float>(K);
                mn.a = go * wn / denom;
                mn.b = -a.stats[3 * map + 0] * kLog2e;
                mn.c = 1.0f / a.stats[3 * map + 1];
                mn.d = a.stats[3 * map + 2];
            }
        }
    };
    const uint32_t lpf_u32 = smem_addr(sh.lp_full), lpe_u32 = smem_addr(sh.lp_empty);
    const uint64_t pol = l2_evict_first_policy();
    const int grp = warp < W ? warp / G : 0, h = warp < W ? warp % G : 0;  // group, warp of the group
    unsigned char* my_slots = s_rdd + static_cast<size_t>(grp) * kGroupBytes;
    const uint32_t slots_u32 = smem_addr(my_slots), bars_u32 = smem_addr(sh.full) + 8 * grp * kGroupBars;
    const int n_my = (q_total > grp) ? (q_total - grp + NG - 1) / NG : 0;  // maps of this group
    const int n_units = n_my * UPM;
    // one lane of a group: request unit u of the group's load sequence into slot `slot` (= u % NS)
    auto request = [&](int u, int slot) {
        const int i = grp + (u / UPM) * NG;  // block-local map
        const size_t off = static_cast<size_t>(m0 + i) * kRDDPixels;
        const float* src = (FUSED && (u % UPM) == 0) ? a.fused + off : a.y_adv + off;
        mbar_arrive_expect_tx(bars_u32 + 8 * slot, kMapBytes);
        bulk_load(slots_u32 + slot * kMapBytes, src, kMapBytes, bars_u32 + 8 * slot, pol);
    };
    // LAZY: the units of the group's map number j: y_adv -> 16 KB slot j & 1, the two heads -> 5 KB slot j % kHeadSlots
    auto request_lazy_p = [&](int j) {
        const int b = j & 1;
        const size_t mp = static_cast<size_t>(m0 + grp + j * NG);
        mbar_arrive_expect_tx(bars_u32 + 8 * b, kMapBytes);
        bulk_load(slots_u32 + b * kMapBytes, a.y_adv + mp * kRDDPixels, kMapBytes, bars_u32 + 8 * b, pol);
    };
    auto request_lazy_h = [&](int j) {
        const int b = j % kHeadSlots;
        const size_t mp = static_cast<size_t>(m0 + grp + j * NG);
        const uint32_t hb = slots_u32 + 2 * kMapBytes + b * kHeadBytes;
        mbar_arrive_expect_tx(bars_u32 + 8 * (2 + b), kHeadBytes);
        bulk_load(hb, a.f_lo + mp * (kRDDLoBytes / 4), kRDDLoBytes, bars_u32 + 8 * (2 + b), pol);
        bulk_load(hb + kRDDLoBytes, a.f_mid + mp * (kRDDMidBytes / 4), kRDDMidBytes, bars_u32 + 8 * (2 + b), pol);
    };
    // The kernel is launched as a PROGRAMMATIC DEPENDENT of the decode launch (hp_decode.cu): its blocks become resident
    // while the decode grid drains.  The first units (inputs the decode does not write) are requested at once; the
    // builder warp alone waits for the decode to complete before it touches the centres, everybody else meets it at the
    // block barrier below.  Nothing is written to global memory before that barrier.
    if (warp < W) {
        if (h == 0 && lane == 0) {
            for (int s = 0; s < kGroupBars; ++s) mbar_init(bars_u32 + 8 * s, 1);
            mbar_init_fence();
            if constexpr (LAZY) {
                for (int j = 0; j < kHeadSlots && j < n_my; ++j) request_lazy_h(j);
                for (int j = 0; j < 2 && j < n_my; ++j) request_lazy_p(j);
            } else {
                for (int u = 0; u < NS && u < n_units; ++u) request(u, u);
            }
        }
    } else {
        griddep_wait();
        if (n_samples > 0) load_sample(sample0);
    }

    // ---- prologue ------------------------------------------------------------------------------------------------------
    // The first sample's label is on every consumer's critical path: it is built by ALL warps, each a band of the map
    // (a single warp needs 3.5-6 us for a sample; the first map of a block used to close 9 us after kernel entry).
    if (warp == W && lane == 0) {
        for (int s = 0; s < 2; ++s) {
            mbar_init(lpf_u32 + 8 * s, 1);
            mbar_init(lpe_u32 + 8 * s, W);
        }
        mbar_init_fence();
    }
    for (int i = threadIdx.x; i < 2 * tmp * tmp + 1; i += blockDim.x) sh.tab[i] = a.tab[i];
    if (threadIdx.x < kFxAccWords) sh.acc[threadIdx.x] = 0ull;
    // everything that does not need the centres happens before the barrier (in the shadow of the decode's tail): the
    // slot tables (from the table in global memory), the band's zeroes
    constexpr int NWARPS = W + 1;
    const int it0 = warp * 32 / NWARPS, it1 = (warp + 1) * 32 / NWARPS;  // this warp's band of the first label: float4 [32 it0, 32 it1)
    RDDSlots<kTileMaxPatch> s6;  // the whole patch (cooperative build, builder)
    rdd_slots_init(s6, a.tab, tmp, a.sdiv, lane, 0, 1);
    RDDSlots<NSL> sl;            // a consumer warp's share of the patch: slot rows h, h + G, ...
    rdd_slots_init(sl, a.tab, tmp, a.sdiv, lane, h, G);
    for (int it = it0; it < it1; ++it) reinterpret_cast<float4*>(lp_base)[it * 32 + lane] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (warp == W && lane < K) {  // waits for the loads issued at kernel entry
        sh.c[0][lane] = cn;
        sh.w[0][lane] = wn;
        if (TASK == RD_BWD) sh.meta[0][lane] = mn;
    }
    __syncthreads();
    if (trace && warp == W && lane == 0) trace[8 + 16 * kRDDTraceMaps * 8] = rdd_now();
    if (n_samples > 0) {
        const int p0 = it0 * 128, p1 = it1 * 128;                            // the band = pixels [p0, p1)
        float* lp = lp_base;
        float4* lp4 = reinterpret_cast<float4*>(lp);
        float* scratch = lp_base + kRDDPixels;  // the second lp slot is free until the builder starts the second sample
        if (trace && threadIdx.x == 0) trace[5] = rdd_now();
        // lane j holds joint j; only the joints whose patch meets the band are visited - ascending, like the builder: the
        // same sums bit for bit
        Centre cl = Centre{0, 0};
        bool meets = false;
        if (lane < K) {
            cl = sh.c[0][lane];
            meets = (cl.y + tmp + 1) * ow > p0 && (cl.y - tmp) * ow < p1;
        }
        unsigned todo = __ballot_sync(0xffffffffu, meets);
        while (todo != 0u) {
            const int j = __ffs(todo) - 1;
            todo &= todo - 1u;
            Centre cj;
            cj.x = __shfl_sync(0xffffffffu, cl.x, j);
            cj.y = __shfl_sync(0xffffffffu, cl.y, j);
            int off[kTileMaxPatch];
            float v[kTileMaxPatch];
#pragma unroll
            for (int k = 0; k < kTileMaxPatch; ++k) {
                const int x = cj.x + s6.dx[k], y = cj.y + s6.dy[k];
                const int o = y * ow + x;
                const bool in = static_cast<unsigned>(x) < static_cast<unsigned>(ow) && o >= p0 && o < p1;
                off[k] = in ? o : -1;
                v[k] = in ? lp[o] : 0.0f;
            }
#pragma unroll
            for (int k = 0; k < kTileMaxPatch; ++k)
                if (off[k] >= 0) lp[off[k]] = v[k] + s6.t[k];
            __syncwarp();
        }
        if (trace && threadIdx.x == 0) trace[6] = rdd_now();
        for (int it = it0; it < it1; ++it) {
            const float4 v = clip01_4(lp4[it * 32 + lane]);
            lp4[it * 32 + lane] = v;
            if (!FUSED && TASK == RD_FWD) {
                scratch[it * 32 + lane] = rdd_term_u(v, eps);
                scratch[1024 + it * 32 + lane] = rdd_term_ulg(v, eps);
            }
        }
    }
    if (trace && threadIdx.x == 0) trace[7] = rdd_now();
    __syncthreads();
    if (trace && threadIdx.x == 0) trace[2] = rdd_now();

    if (warp == W) {
        // =================================== builder warp: lp of every sample of the block, one ahead ==================
        // (the second sample's loads only now: queued behind the first bulk copies they take microseconds, and a wait
        // for them inside the cooperative build would hold every warp at its barrier)
        if (n_samples > 1) load_sample(sample0 + 1);
        const RDDSlots<kTileMaxPatch>& sl = s6;
        if (n_samples > 0) {  // the first sample: built by all warps above; its constants in the canonical order
            if (!FUSED && TASK == RD_FWD) {
                const float* scratch = lp_base + kRDDPixels;
                float su = 0.f, sulg = 0.f;
#pragma unroll 8
                for (int it = 0; it < 32; ++it) {
                    su += scratch[it * 32 + lane];
                    sulg += scratch[1024 + it * 32 + lane];
                }
                const float rr = warp_sum3_scattered(su, sulg, 0.0f, lane);
                if (lane == 0) sh.cu[0] = rr;
                if (lane == 8) sh.culg[0] = rr;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(lpf_u32);
            if (trace && lane == 0) trace[8 + 16 * kRDDTraceMaps * 8 + 1] = rdd_now();
        }
        for (int r = 1; r < n_samples; ++r) {
            const int slot = r & 1;
            const Centre cc = cn;
            const float ww = wn;
            const RDSMeta mm = mn;
            if (r + 1 < n_samples) load_sample(sample0 + r + 1);
            if (r >= 2) mbar_wait_backoff(lpe_u32 + 8 * slot, static_cast<uint32_t>((r >> 1) - 1) & 1u, 100);
            if (trace && lane == 0 && r < 8) trace[8 + 16 * kRDDTraceMaps * 8 + 2 * r] = rdd_now();
            float* lp = lp_base + slot * kRDDPixels;
            float4* lp4 = reinterpret_cast<float4*>(lp);
#pragma unroll 8
            for (int it = 0; it < 32; ++it) lp4[it * 32 + lane] = make_float4(0.f, 0.f, 0.f, 0.f);
            __syncwarp();
            for (int j = 0; j < K; ++j) {  // joints ascending: the summation order of every pixel is fixed
                const int cx = __shfl_sync(0xffffffffu, cc.x, j), cy = __shfl_sync(0xffffffffu, cc.y, j);
                // the patch pixels of one joint are distinct: all loads, then all stores (no false dependencies)
                int off[kTileMaxPatch];
                float v[kTileMaxPatch];
#pragma unroll
                for (int k = 0; k < kTileMaxPatch; ++k) {
                    const int x = cx + sl.dx[k], y = cy + sl.dy[k];
                    const bool in = static_cast<unsigned>(x) < static_cast<unsigned>(ow) && static_cast<unsigned>(y) < static_cast<unsigned>(oh);
                    off[k] = in ? y * ow + x : -1;
                    v[k] = in ? lp[off[k]] : 0.0f;
                }
#pragma unroll
                for (int k = 0; k < kTileMaxPatch; ++k)
                    if (off[k] >= 0) lp[off[k]] = v[k] + sl.t[k];
                __syncwarp();
            }
            float su = 0.f, sulg = 0.f;
#pragma unroll 4
            for (int it = 0; it < 32; ++it) {
                const float4 v = clip01_4(lp4[it * 32 + lane]);
                lp4[it * 32 + lane] = v;
                if (!FUSED && TASK == RD_FWD) {
                    su += rdd_term_u(v, eps);
                    sulg += rdd_term_ulg(v, eps);
                }
            }
            if (!FUSED && TASK == RD_FWD) {
                const float rr = warp_sum3_scattered(su, sulg, 0.0f, lane);
                if (lane == 0) sh.cu[slot] = rr;
                if (lane == 8) sh.culg[slot] = rr;
            }
            if (lane < K) {
                sh.c[slot][lane] = cc;
                sh.w[slot][lane] = ww;
                if (TASK == RD_BWD) sh.meta[slot][lane] = mm;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(lpf_u32 + 8 * slot);
            if (trace && lane == 0 && r < 8) trace[8 + 16 * kRDDTraceMaps * 8 + 2 * r + 1] = rdd_now();
        }
    } else {
        // =================================== consumer warps: G warps per map ===========================================
        const float2 l2e = make_float2(kLog2e, kLog2e);
        const int bar_id = 1 + grp;
        uint32_t par_mask = 0;        // bit s: parity of slot s's next completion
        int slot_c = 0;               // slot of the next unit to consume
        int u_load = (n_units < NS) ? n_units : NS;   // next unit to request (its slot is the one consumed NS units before)
        auto wait_unit = [&]() {      // the next unit of the sequence has landed; returns its slot
            const int s = slot_c;
            mbar_wait(bars_u32 + 8 * s, (par_mask >> s) & 1u);
            par_mask ^= 1u << s;
            slot_c = (s + 1 == NS) ? 0 : s + 1;
            return s;
        };
        // LAZY: heads and y_adv of the group's map jj have landed; their byte offsets inside the group's slots
        auto wait_lazy = [&](int j, uint32_t& f_off, uint32_t& p_off) {
            const int b = j % kHeadSlots;
            const uint32_t parity = static_cast<uint32_t>(j / kHeadSlots) & 1u;
            mbar_wait(bars_u32 + 8 * (2 + b), parity);  // the heads; y_adv is awaited (wait_lazy_p) after the fused blocks are built
            f_off = 2 * kMapBytes + b * kHeadBytes;
            p_off = (j & 1) * kMapBytes;
        };
        auto wait_lazy_p = [&](int j) { mbar_wait(bars_u32 + 8 * (j & 1), static_cast<uint32_t>(j >> 1) & 1u); };
        int r_done = 0;  // samples [0, r_done) of the block have been released by this warp
        int k = m0 + grp - sample0 * K, r = 0;  // joint and block-relative sample of the current map
        // wait for a sample's lp; the samples before it are released in order, each only after ITS lp has been seen
        // complete (so that the empty-barrier of a slot never receives arrivals of two phases at once)
        auto advance_to = [&](int r_want) {
            while (r_done < r_want) {
                mbar_wait_backoff(lpf_u32 + 8 * (r_done & 1), static_cast<uint32_t>(r_done >> 1) & 1u, 50);
                __syncwarp();
                if (lane == 0) mbar_arrive(lpe_u32 + 8 * (r_done & 1));
                ++r_done;
            }
            if (r_want < n_samples) mbar_wait(lpf_u32 + 8 * (r_want & 1), static_cast<uint32_t>(r_want >> 1) & 1u);
        };
        // LAZY: this lane's two 4x4 blocks of the fused map - column block bn, block rows bm, bm + 1 (the walk of
        // fuse_block_kernel<4, 2, ., 2>: warp h of the group covers output rows 16 h .. 16 h + 15)
        const int bn = lane & 15, bm = (h * 2 + (lane >> 4)) * 2;
        FxReg fx{0, 0, 0, 0, 0};
        int jj = 0;  // the group's map counter
        for (int i = grp; i < q_total; i += NG, ++jj) {
            while (k >= K) {
                k -= K;
                ++r;
            }
            const int map = m0 + i, slot = r & 1, xb = jj & 1;
            const bool leader = (jj % G) == h;
            const float* lp = lp_base + slot * kRDDPixels;
            const float4* LP4 = reinterpret_cast<const float4*>(lp);
            float* xch = sh.xch[xb][warp];
            int freed0;  // first slot freed by this map (UPM consecutive slots)
            unsigned long long* wt = (trace && lane == 0 && jj < kRDDTraceMaps && warp < 16) ? trace + 8 + (warp * kRDDTraceMaps + jj) * 8 : nullptr;
            if (wt) wt[0] = rdd_now();
            if (TASK == RD_BWD) {
                // ---- backward: d/dp = coef (softmax(p) - u / S),  u = g / M + eps   -----------------
                uint32_t f_off = 0, p_off = 0;
                if constexpr (LAZY) {
                    wait_lazy(jj, f_off, p_off);
                    freed0 = 0;
                } else {
                    const int sf = FUSED ? wait_unit() : 0, sp = wait_unit();
                    freed0 = FUSED ? sf : sp;
                    f_off = sf * kMapBytes;
                    p_off = sp * kMapBytes;
                }
                const float4* P4 = reinterpret_cast<const float4*>(my_slots + p_off);
                const float* P = reinterpret_cast<const float*>(P4);
                const float4* F4 = reinterpret_cast<const float4*>(my_slots + f_off);  // FUSED only
                const float* F = reinterpret_cast<const float*>(F4);                            // (LAZY: the heads, 16x16 then 32x32)
                if (wt) wt[1] = rdd_now();
                float4 fv[LAZY ? 8 : 1];
                if constexpr (LAZY)  // before the label is awaited: the fused values of this lane's blocks
                {
                    rdd_fused_blocks(slots_u32 + f_off, slots_u32 + f_off + kRDDLoBytes, a.a_lo, a.a_mid, bn, bm, fv);
                    wait_lazy_p(jj);
                }
                advance_to(r);
                if (wt) wt[2] = rdd_now();
                const Centre ck = sh.c[slot][k];
                const RDSMeta meta = sh.meta[slot][k];
                const float coef = meta.a, lb = meta.b, invS = meta.c, M = meta.d;
                const float invM = (M == 1.0f) ? 1.0f : __frcp_rn(M);
                const float k1 = -coef * (invM * invS), k0 = -coef * (eps * invS);  // -coef u / S = k1 g + k0
                float* gout = a.grad_in + static_cast<size_t>(map) * kRDDPixels;
                // the own patch (this warp's slots): exact values, written after the generic pass
                float gk[NSL];
                int poff[NSL];
#pragma unroll
                for (int kk = 0; kk < NSL; ++kk) {
                    const int x = ck.x + sl.dx[kk], y = ck.y + sl.dy[kk];
                    const bool in = static_cast<unsigned>(x) < static_cast<unsigned>(ow) && static_cast<unsigned>(y) < static_cast<unsigned>(oh);
                    const int off = in ? y * ow + x : 0;
                    float g = rdd_patch_value(a.variant, lp[off], sl.t[kk], k, K, x, y, sh.tab, tmp, sh.c[slot]);
                    if (FUSED) {
                        const float fval = LAZY ? rdd_fused_at(F, F + kRDDLoBytes / 4, a.a_lo, a.a_mid, in ? x : 0, in ? y : 0) : F[off];
                        g = clip01(__fsub_rn(__fadd_rn(g, fval), __fmul_rn(sl.t[kk], 100.0f)));
                    }
                    gk[kk] = fmaf(coef, ex2_approx(fmaf(P[off], kLog2e, lb)), fmaf(g, k1, k0));
                    poff[kk] = in ? off : -1;
                }
                if (wt) wt[3] = rdd_now();
                const float2 lb2 = make_float2(lb, lb), k12 = make_float2(k1, k1), k02 = make_float2(k0, k0), c2 = make_float2(coef, coef);
                float4* out4 = reinterpret_cast<float4*>(gout);
                auto grad4 = [&](int i4, float4 g) {  // float4 number i4 of the map: gradient from the prediction and g
                    const float4 v = P4[i4];
                    const float2 a0 = __ffma2_rn(make_float2(v.x, v.y), l2e, lb2), a1 = __ffma2_rn(make_float2(v.z, v.w), l2e, lb2);
                    const float2 q0 = __ffma2_rn(make_float2(g.x, g.y), k12, k02), q1 = __ffma2_rn(make_float2(g.z, g.w), k12, k02);
                    const float2 r0 = __ffma2_rn(c2, make_float2(ex2_approx(a0.x), ex2_approx(a0.y)), q0);
                    const float2 r1 = __ffma2_rn(c2, make_float2(ex2_approx(a1.x), ex2_approx(a1.y)), q1);
                    stg_stream4(out4 + i4, make_float4(r0.x, r0.y, r1.x, r1.y));
                };
                if constexpr (LAZY) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) {  // the lane's blocks: row 4 bm + q, columns 4 bn .. 4 bn + 3
                        const int i4 = (4 * bm + q) * kRDDRow4 + bn;
                        grad4(i4, clip01_4(add4(LP4[i4], fv[q])));
                    }
                } else {
#pragma unroll
                    for (int it = h * IT; it < (h + 1) * IT; ++it) {
                        float4 g = LP4[it * 32 + lane];
                        if (FUSED) g = clip01_4(add4(g, F4[it * 32 + lane]));
                        grad4(it * 32 + lane, g);
                    }
                }
                if (wt) wt[4] = rdd_now();
                // every warp of the group has issued its stores of the map (and read the slots out): the patch pixels are
                // overwritten with their exact values (ordered behind the generic stores by the barrier)
                group_barrier(bar_id, 32 * G);
#pragma unroll
                for (int kk = 0; kk < NSL; ++kk)
                    if (poff[kk] >= 0) gout[poff[kk]] = gk[kk];
                if (leader && lane == 0) {
                    if constexpr (LAZY) {
                        if (jj + 2 < n_my) {  // (backward: two head slots)
                            request_lazy_h(jj + 2);
                            request_lazy_p(jj + 2);
                        }
                    } else {
                        int s = freed0;
                        for (int c = 0; c < UPM; ++c) {
                            if (u_load + c < n_units) request(u_load + c, s);
                            s = (s + 1 == NS) ? 0 : s + 1;
                        }
                    }
                }
                if (wt) wt[5] = rdd_now();
                u_load += UPM;
            } else if (!FUSED) {
                const int sp = wait_unit();
                freed0 = sp;
                const float4* P4 = reinterpret_cast<const float4*>(my_slots + sp * kMapBytes);
                const float* P = reinterpret_cast<const float*>(P4);
                // ---- the softmax reference: the maximum of the map's first iteration (no pass over the map; every warp of
                //      the group reads the same 32 float4) -------------------------------------------------------------------
                float Mp = warp_max_f32(max4(P4[lane]));
                const float ms = (Mp == -INFINITY) ? 0.0f : Mp;
                const float mb = -ms * kLog2e;
                if (wt) wt[1] = rdd_now();
                advance_to(r);
                if (wt) wt[2] = rdd_now();
                const Centre ck = sh.c[slot][k];
                // M = 1 iff some other centre lies outside the own patch (lp = 1 there, own Gaussian 0)
                bool outside = false;
                if (lane < K) {
                    const Centre cj = sh.c[slot][lane];
                    outside = abs(cj.x - ck.x) > tmp || abs(cj.y - ck.y) > tmp;
                }
                const bool fast = __any_sync(0xffffffffu, outside) || a.variant != HP_RD_X6;  // rd4 and base do not normalise
                // ---- own patch (this warp's slots): exact recipe minus the generic term --------------------------------
                float c_u = 0.f, c_up = 0.f, c_ulg = 0.f;
#pragma unroll
                for (int kk = 0; kk < NSL; ++kk) {
                    const int x = ck.x + sl.dx[kk], y = ck.y + sl.dy[kk];
                    const bool in = static_cast<unsigned>(x) < static_cast<unsigned>(ow) && static_cast<unsigned>(y) < static_cast<unsigned>(oh);
                    const int off = in ? y * ow + x : 0;
                    const float lpv = lp[off], pk = P[off];
                    const float gex = rdd_patch_value(a.variant, lpv, sl.t[kk], k, K, x, y, sh.tab, tmp, sh.c[slot]);
                    const float du = in ? gex - lpv : 0.0f;
                    const float dl = ulg2(gex + eps) - ulg2(lpv + eps);
                    c_u += du;
                    c_up = in ? fmaf(du, fmaf(pk, kLog2e, mb), c_up) : c_up;
                    c_ulg += in ? dl : 0.0f;
                }
                if (wt) wt[3] = rdd_now();
                // ---- pass B: softmax sum against the reference, sum u a with u = lp + eps, a = (p - ref) log2 e ---------------
                const float2 mb2 = make_float2(mb, mb), e2 = make_float2(eps, eps);
                float2 s2 = make_float2(0.f, 0.f), slp2 = make_float2(0.f, 0.f), slq2 = make_float2(0.f, 0.f);
#pragma unroll
                for (int it = h * IT; it < (h + 1) * IT; ++it) {
                    const float4 v = P4[it * 32 + lane];
                    const float4 l = LP4[it * 32 + lane];
                    const float2 lo = make_float2(v.x, v.y), hi = make_float2(v.z, v.w);
                    const float2 a0 = __ffma2_rn(lo, l2e, mb2), a1 = __ffma2_rn(hi, l2e, mb2);
                    s2 = __fadd2_rn(s2, __fadd2_rn(make_float2(ex2_approx(a0.x), ex2_approx(a0.y)),
                                                   make_float2(ex2_approx(a1.x), ex2_approx(a1.y))));
                    slp2 = __ffma2_rn(__fadd2_rn(make_float2(l.x, l.y), e2), a0, slp2);
                    slq2 = __ffma2_rn(__fadd2_rn(make_float2(l.z, l.w), e2), a1, slq2);
                }
                if (wt) wt[4] = rdd_now();
                {
                    const float rr = warp_sum3_scattered(s2.x + s2.y, slp2.x + slp2.y + slq2.x + slq2.y, 0.0f, lane);
                    const float r2 = warp_sum3_scattered(c_u, c_up, c_ulg, lane);
                    // lanes 0 / 8 / 16 hold the three sums of each reduction
                    if ((lane & 7) == 0 && lane < 16) xch[lane >> 3] = rr;
                    if ((lane & 7) == 0 && lane < 24) xch[2 + (lane >> 3)] = r2;
                }
                group_barrier(bar_id, 32 * G);
                if (leader) {
                    float Sexp = 0.f, Sua = 0.f, cu = 0.f, culg = 0.f;
#pragma unroll
                    for (int g = 0; g < G; ++g) {
                        const float* x = sh.xch[xb][grp * G + g];
                        Sexp += x[0];
                        Sua += x[1] + x[3];
                        cu += x[2];
                        culg += x[4];
                    }
                    float Su = sh.cu[slot] + cu, Sulg = sh.culg[slot] + culg, M = 1.0f;
                    if (!fast) rdd_exact_map(P4, LP4, sh.tab, tmp, ck, a.wdiv, eps, mb, lane, Su, Sua, Sulg, M);  // the slot is still intact
                    if (Sexp == INFINITY || Sexp == 0.0f) {
                        // the sampled reference lies too far below the map's maximum (or is -inf): the exact two-pass sum,
                        // and sum u a re-based to the true maximum
                        float mx, sx;
                        rdd_softmax_exact(P4, 0, 32, lane, mx, sx);
                        const float msx = (mx == -INFINITY) ? 0.0f : mx;
                        Sua = fmaf((ms - msx) * kLog2e, Su, Sua);
                        Mp = mx;
                        Sexp = sx;
                    }
                    __syncwarp();
                    if (lane == 0) {
                        if (u_load < n_units) request(u_load, freed0);
                        if (wt) wt[5] = rdd_now();
                        // ---- closure: L = (sum u ln u - sum u p)/S - ln S + lse   (loss.py:145-158), in log2 units against
                        //      the softmax reference: L = ln 2 ((sum u lg u - sum u a)/S - lg S + lg sum 2^a) ---------------------
                        const float lg_se = lg2_approx(Sexp);
                        const float lse = fmaf(lg_se, kLn2, (Mp == -INFINITY) ? 0.0f : Mp);
                        const float L = kLn2 * (__fdividef(Sulg - Sua, Su) - lg2_approx(Su) + lg_se);
                        const float Lw = L * sh.w[slot][k];
                        a.per_map[map] = Lw;
                        if (a.mean) fx_reg_add(fx, Lw);
                        a.stats[3 * map + 0] = lse;
                        a.stats[3 * map + 1] = Su;
                        a.stats[3 * map + 2] = M;
                    }
                }
                u_load += UPM;
            } else {
                uint32_t f_off = 0, p_off = 0;
                if constexpr (LAZY) {
                    wait_lazy(jj, f_off, p_off);
                    freed0 = 0;
                } else {
                    // the fused map now; y_adv is awaited behind pass 1, which does not read it.  g stays in registers, so the
                    // fused map's slot is free once every warp of the group has been through pass 1: the NEXT map's y_adv is
                    // requested into it there (not when this map closes), a whole pass 2 + closure + patch + pass 1 ahead of
                    // its use (timeline before: 1.0 us of a map's 3.1 us were the wait for y_adv)
                    const int sf = wait_unit(), sp = slot_c;
                    freed0 = sf;
                    f_off = sf * kMapBytes;
                    p_off = sp * kMapBytes;
                }
                const float4* P4 = reinterpret_cast<const float4*>(my_slots + p_off);
                const float* P = reinterpret_cast<const float*>(P4);
                float4* F4 = reinterpret_cast<float4*>(my_slots + f_off);
                float* F = reinterpret_cast<float*>(F4);  // (LAZY: the heads, 16x16 then 32x32; read only)
                if (wt) wt[1] = rdd_now();
                float4 fv[LAZY ? 8 : IT];  // g of this lane's float4s (LAZY: first the fused values of its two blocks)
                if constexpr (LAZY)  // before the label is awaited: the fused values of this lane's blocks
                {
                    rdd_fused_blocks(slots_u32 + f_off, slots_u32 + f_off + kRDDLoBytes, a.a_lo, a.a_mid, bn, bm, fv);
                    wait_lazy_p(jj);
                }
                advance_to(r);
                if (wt) wt[2] = rdd_now();
                const Centre ck = sh.c[slot][k];
                // ---- own patch (this warp's slots): exact un-normalised values into registers, -inf into the fused map
                //      (the passes then see g = 0 there) -------------------------------------------------------------------
                float gex[NSL], pk[NSL];
                int poff[NSL];
                float mg = -INFINITY;
#pragma unroll
                for (int kk = 0; kk < NSL; ++kk) {
                    const int x = ck.x + sl.dx[kk], y = ck.y + sl.dy[kk];
                    const bool in = static_cast<unsigned>(x) < static_cast<unsigned>(ow) && static_cast<unsigned>(y) < static_cast<unsigned>(oh);
                    const int off = in ? y * ow + x : 0;
                    const float lpv = lp[off];
                    const float fval = LAZY ? rdd_fused_at(F, F + kRDDLoBytes / 4, a.a_lo, a.a_mid, in ? x : 0, in ? y : 0) : F[off];
                    pk[kk] = LAZY ? P[off] : 0.0f;  // (pre-fused: read behind pass 1, once y_adv has been awaited)
                    float g = clip01(__fsub_rn(lpv, __fmul_rn(sl.t[kk], 10.0f)));
                    g = clip01(__fsub_rn(__fadd_rn(g, fval), __fmul_rn(sl.t[kk], 100.0f)));
                    gex[kk] = g;
                    poff[kk] = in ? off : -1;
                    mg = in ? fmaxf(mg, g) : mg;
                }
                if constexpr (LAZY) {
                    // ---- pass 1 in registers: g = clip(lp + f) of the lane's blocks, 0 on the own patch (its exact values are
                    //      in gex); nothing is written to the slot, so no barrier separates the patch step from this pass
                    if (wt) wt[3] = rdd_now();
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const int Y = 4 * bm + q;
                        float4 g = clip01_4(add4(LP4[Y * kRDDRow4 + bn], fv[q]));
                        if (static_cast<unsigned>(Y - ck.y + tmp) <= static_cast<unsigned>(2 * tmp)) {
                            const int dx0 = 4 * bn - ck.x + tmp;  // column c of the float4 lies in the patch iff 0 <= dx0 + c <= 2 tmp
                            g.x = (static_cast<unsigned>(dx0 + 0) <= static_cast<unsigned>(2 * tmp)) ? 0.0f : g.x;
                            g.y = (static_cast<unsigned>(dx0 + 1) <= static_cast<unsigned>(2 * tmp)) ? 0.0f : g.y;
                            g.z = (static_cast<unsigned>(dx0 + 2) <= static_cast<unsigned>(2 * tmp)) ? 0.0f : g.z;
                            g.w = (static_cast<unsigned>(dx0 + 3) <= static_cast<unsigned>(2 * tmp)) ? 0.0f : g.w;
                        }
                        fv[q] = g;
                        mg = fmaxf(mg, max4(g));
                    }
                } else {
#pragma unroll
                    for (int kk = 0; kk < NSL; ++kk)
                        if (poff[kk] >= 0) F[poff[kk]] = -INFINITY;   // a patch pixel belongs to exactly one lane of the group
                    group_barrier(bar_id, 32 * G);
                    if (wt) wt[3] = rdd_now();
                    // ---- pass 1: g = clip(lp + f) written over f, its maximum; the softmax reference is sampled ------------
#pragma unroll
                    for (int it = h * IT; it < (h + 1) * IT; ++it) {
                        const float4 f = F4[it * 32 + lane], l = LP4[it * 32 + lane];
                        const float4 g = clip01_4(add4(l, f));
                        fv[LAZY ? 0 : it - h * IT] = g;
                        mg = fmaxf(mg, max4(g));
                    }
                    wait_unit();  // y_adv (slot sp)
#pragma unroll
                    for (int kk = 0; kk < NSL; ++kk) pk[kk] = P[poff[kk] >= 0 ? poff[kk] : 0];
                }
                {
                    const float wmg = warp_max_f32(mg);
                    if (lane == 0) sh.xmx[xb][warp] = wmg;
                }
                float Mp = warp_max_f32(max4(P4[lane]));  // the softmax reference: sampled from the map's first iteration
                if constexpr (!LAZY) fence_proxy_async_smem();  // this lane's -inf marks in the slot precede the copy engine's next write
                group_barrier(bar_id, 32 * G);
                if constexpr (!LAZY) {  // every warp of the group is through pass 1: the fused map's slot takes the next map's y_adv
                    if (leader && lane == 0 && u_load < n_units) request(u_load, freed0);
                }
                if constexpr (LAZY && kHeadSlots == 1) {  // every warp of the group has read the heads (blocks and patch): the next map's
                    if (leader && lane == 0 && jj + 1 < n_my) request_lazy_h(jj + 1);
                }
                float M = -INFINITY;
#pragma unroll
                for (int g = 0; g < G; ++g) M = fmaxf(M, sh.xmx[xb][grp * G + g]);
                const float invM = (M == 1.0f) ? 1.0f : __frcp_rn(M);
                // ---- pass 2: the sums of u = g / M + eps ----------------------------------------------------------------
                const float ms = (Mp == -INFINITY) ? 0.0f : Mp;
                const float mb = -ms * kLog2e;
                const float2 mb2 = make_float2(mb, mb), im2 = make_float2(invM, invM), e2 = make_float2(eps, eps);
                const float2 z2 = make_float2(0.f, 0.f);
                RDDFusedAcc A{z2, z2, z2, z2, z2, z2};
                if constexpr (LAZY) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) rdd_fused_acc(A, P4[(4 * bm + q) * kRDDRow4 + bn], fv[q], l2e, mb2, im2, e2);
                } else {
#pragma unroll
                    for (int it = h * IT; it < (h + 1) * IT; ++it) rdd_fused_acc(A, P4[it * 32 + lane], fv[LAZY ? 0 : it - h * IT], l2e, mb2, im2, e2);
                }
                const float2 s2 = A.s2, su2 = A.su2, sup2 = A.sup2, suq2 = A.suq2, sulg2 = A.sulg2, sulh2 = A.sulh2;
                if (wt) wt[4] = rdd_now();
                // this warp's patch pixels went through the passes as g = 0 (u = eps): replace by the exact values
                float c_u = 0.f, c_up = 0.f, c_ulg = 0.f;
                const float u_bg = fmaf(0.0f, invM, eps), bg_ulg = ulg2(u_bg);
#pragma unroll
                for (int kk = 0; kk < NSL; ++kk) {
                    const float uex = fmaf(gex[kk], invM, eps);
                    const float du = uex - u_bg, dl = ulg2(uex) - bg_ulg;
                    const bool in = poff[kk] >= 0;
                    c_u += in ? du : 0.0f;
                    c_up = in ? fmaf(du, fmaf(pk[kk], kLog2e, mb), c_up) : c_up;
                    c_ulg += in ? dl : 0.0f;
                }
                {
                    const float rr = warp_sum3_scattered(s2.x + s2.y, su2.x + su2.y + c_u, sup2.x + sup2.y + suq2.x + suq2.y + c_up, lane);
                    float t = sulg2.x + sulg2.y + sulh2.x + sulh2.y + c_ulg;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
                    if ((lane & 7) == 0 && lane < 24) xch[lane >> 3] = rr;
                    if (lane == 24) xch[3] = t;
                }
                fence_proxy_async_smem();  // this lane's writes to the slot precede the copy engine's next write
                group_barrier(bar_id, 32 * G);
                if (leader) {
                    float Sexp = 0.f, Su = 0.f, Sup = 0.f, Sulg = 0.f;
#pragma unroll
                    for (int g = 0; g < G; ++g) {
                        const float* x = sh.xch[xb][grp * G + g];
                        Sexp += x[0];
                        Su += x[1];
                        Sup += x[2];
                        Sulg += x[3];
                    }
                    if (Sexp == INFINITY || Sexp == 0.0f) {  // see the branch without a fused map; the slot is still intact
                        float mx, sx;
                        rdd_softmax_exact(P4, 0, 32, lane, mx, sx);
                        const float msx = (mx == -INFINITY) ? 0.0f : mx;
                        Sup = fmaf((ms - msx) * kLog2e, Su, Sup);
                        Mp = mx;
                        Sexp = sx;
                    }
                    __syncwarp();
                    if (lane == 0) {
                        if constexpr (LAZY) {
                            if (jj + 2 < n_my) request_lazy_p(jj + 2);
                            if (kHeadSlots == 2 && jj + 2 < n_my) request_lazy_h(jj + 2);
                        } else {  // (the next map's y_adv, unit u_load, went out behind pass 1) the fused map after it -> y_adv's slot
                            if (u_load + 1 < n_units) request(u_load + 1, (freed0 + 1 == NS) ? 0 : freed0 + 1);
                        }
                        if (wt) wt[5] = rdd_now();
                        const float lg_se = lg2_approx(Sexp);
                        const float lse = fmaf(lg_se, kLn2, (Mp == -INFINITY) ? 0.0f : Mp);
                        const float L = kLn2 * (__fdividef(Sulg - Sup, Su) - lg2_approx(Su) + lg_se);
                        const float Lw = L * sh.w[slot][k];
                        a.per_map[map] = Lw;
                        if (a.mean) fx_reg_add(fx, Lw);
                        a.stats[3 * map + 0] = lse;
                        a.stats[3 * map + 1] = Su;
                        a.stats[3 * map + 2] = M;
                    }
                }
                u_load += UPM;
            }
            if (wt) wt[6] = rdd_now();
            k += NG;
        }
        advance_to(n_samples);  // release the remaining samples (keeps the arrival counts of the empty barriers whole)
        if (TASK == RD_FWD && lane == 0 && a.mean) fx_reg_flush(fx, sh.acc);
    }

    // ---- epilogue: block sum -> workspace, the last block finalises 'mean' / the per-sample means ----------------------
    if (trace) {
        __syncthreads();
        if (threadIdx.x == 0) trace[1] = rdd_now();
    }
    if (TASK != RD_FWD || (a.mean == nullptr && a.per_sample == nullptr)) return;
    __syncthreads();
    if (a.mean && threadIdx.x < kFxAccWords && sh.acc[threadIdx.x] != 0ull) atomicAdd(&a.ws->acc[threadIdx.x], sh.acc[threadIdx.x]);
    if (last_block_arrives(&a.ws->counter, gridDim.x)) {
        if (a.per_sample) per_sample_means(a.per_map, a.B, K, a.per_sample, threadIdx.x, 32 * (W + 1));
        if (a.mean && threadIdx.x == 0) *a.mean = fx_mean_from_workspace(a.ws->acc, n_maps);
        if (threadIdx.x == 0) a.ws->counter = 0;
    }
}

template <bool FUSED, int GW, int TASK, bool LAZY = false>
static int launch_rdd_shape(const RDArgs& a, int sms, cudaStream_t stream, const char* who) {
    using Shape = RDDShape<FUSED, GW, LAZY ? kRDDLazyGroups : 0>;
    constexpr size_t smem = LAZY ? Shape::kLazySmem : Shape::kSmem;
    static bool attr_done_dev[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    bool& attr_done = attr_done_dev[dev & 63];
    if (!attr_done) {
        const cudaError_t e = cudaFuncSetAttribute(regdisp_dense_kernel<FUSED, GW, TASK, LAZY>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                   static_cast<int>(smem));
        if (e != cudaSuccess) return fail(static_cast<int>(e), "%s: %s", who, cudaGetErrorString(e));
        attr_done = true;
    }
    const int n_maps = a.B * a.K;
    int grid = sms;
    if (const char* e = getenv("HP_RD_GRID")) {  // tests: few blocks -> long map ranges spanning many samples
        const int g = atoi(e);
        if (g > 0) grid = g;
    }
    if (grid > n_maps) grid = n_maps;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(32 * (Shape::W + 1));
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;  // behind the decode launch of hp_regdisp_fwd
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (TASK == RD_FWD) ? 1 : 0;
    const cudaError_t e = cudaLaunchKernelEx(&cfg, regdisp_dense_kernel<FUSED, GW, TASK, LAZY>, a);
    if (e != cudaSuccess) return fail(static_cast<int>(e), "%s: %s", who, cudaGetErrorString(e));
    return launch_status(who);
}

// profiling only (hp_debug_regdisp_trace)
static unsigned long long* g_rdd_trace = nullptr;
static size_t g_rdd_trace_words = 0;

// forward / backward, mode 'max', x6 / rd4 recipes on 4096-pixel maps; returns 1 when the case is not covered (the caller takes
// the register-slice kernel), 0 when launched.  HP_RD_DENSE=0 keeps the old kernel (comparison runs).
template <int TASK>
static int launch_regdisp_dense(RDArgs a, cudaStream_t stream, const char* who) {
    static const bool on = []() {
        const char* e = getenv("HP_RD_DENSE");
        return !(e && e[0] == '0');
    }();
    if (!on) return 1;
    if (a.mode != HP_MODE_MAX || (a.variant != HP_RD_X6 && a.variant != HP_RD_RD4 && a.variant != HP_RD_BASE)) return 1;
    if (a.oh * a.ow != kRDDPixels || a.ow % 4 != 0 || a.K > kRDDMaxK || a.tmp > 6) return 1;
    if ((2 * a.tmp + 1) * (2 * a.tmp + 1) > 32 * kTileMaxPatch) return 1;
    static int sms = 0;
    if (sms == 0) {
        sms = hp_device_sm_count();
        if (sms <= 0) sms = 148;
    }
    a.wdiv = FastDiv(static_cast<uint32_t>(a.ow));
    a.sdiv = FastDiv(static_cast<uint32_t>(2 * a.tmp + 1));
    a.trace = (g_rdd_trace && g_rdd_trace_words >= static_cast<size_t>(sms) * kRDDTraceBlockWords) ? g_rdd_trace : nullptr;
    if (a.f_lo != nullptr) {  // the fused map built in the kernel from its two heads (hp_regdisp_fwd_heads / _bwd_heads): x6 only
        if (a.variant != HP_RD_X6 || a.f_mid == nullptr || a.fused != nullptr || a.ow != 64 || a.oh != 64) return 1;
        return launch_rdd_shape<true, 4, TASK, true>(a, sms, stream, who);
    }
    const bool fused = a.fused != nullptr && a.variant == HP_RD_X6;
    // warps per map with a fused map: HP_RDD_G=2|4 overrides the default (comparison runs)
    static const int g_env = []() {
        const char* e = getenv("HP_RDD_G");
        return e ? atoi(e) : 0;
    }();
    if (fused) return g_env == 2 ? launch_rdd_shape<true, 2, TASK>(a, sms, stream, who) : launch_rdd_shape<true, 4, TASK>(a, sms, stream, who);
    return launch_rdd_shape<false, 2, TASK>(a, sms, stream, who);  // (4 warps per map measured slower: 82 vs 75 us with the decode)
}

}  // namespace hp
