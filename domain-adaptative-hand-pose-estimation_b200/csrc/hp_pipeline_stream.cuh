// hp_pipeline_stream.cuh - the fast shape of the fused gen+loss+decode+PCK kernel.
//
// One WARP per map, no block barrier on the data path.  The first ncu capture of the block-per-map
// kernel (profiles/r1_pipeline_v1.md) showed 43 issued instructions per element at 13 % of HBM peak:
// the kernel was issue-bound, not memory-bound.  At 6.45 TB/s a B200 SM has ~730 cycles (~2900 issue
// slots) per 16 KB map, i.e. ~22 instructions per element, so this kernel is written to an
// instruction budget:
//   * hot loop sees no target at all: per element  max (FMNMX3), ==max index scan (FSETP+SEL),
//     packed FFMA2 for (p*log2e - m*log2e), MUFU.EX2, packed FADD2 for sum exp and sum p, packed
//     FFMA2 for sum p^2  -> ~6.5 instructions per element;
//   * the Gaussian target only exists on <= (2*tmp+1)^2 = 169 pixels: each lane re-reads its <= 6
//     patch pixels up front (L2 hits on lines the tile loads fetch anyway) and the target terms
//     sum_patch{u*p, u*log u, u, p, t*(t-2p)} are added to the closed-form background terms;
//   * tiles of 8 x 128-bit loads per lane (4 KB per warp) are double-buffered in registers, so every
//     resident warp always has 4 KB in flight (16 warps/SM -> 64 KB/SM, Little's law needs ~26 KB);
//   * per-map scalar work (centre, PCK distance, KL closure) takes fp32/integer fast paths that are
//     exactly equivalent to the reference's float64 arithmetic, and falls back to float64 only when a
//     value is within 1e-4 of a decision boundary.
#pragma once
#include "hp_common.cuh"

namespace hp {

constexpr int kStreamWarps = 4;          // warps (= maps) per block
constexpr int kStreamMaxPatch = 6;       // patch pixels per lane: (2*tmp+1)^2 <= 192

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

struct StreamArgs {
    const float* pred;
    const double* joints;
    const float* vis;
    int n_maps, K, H, W, HW, ntiles;
    FastDiv wdiv, sdiv;       // by W ; by patch side (2*tmp+1)
    double sx, sy, inv_sx, inv_sy;
    int pow2_stride;          // joint / stride == joint * inv_stride exactly
    int tmp;
    const float* tab;
    float eps, eps_log_eps;   // eps*ln(eps) (0 when eps == 0)
    double thr;
    float thr2_lo, thr2_hi, inv_nx, inv_ny;   // fp32 pre-test of the PCK distance
    float* pred_xy;
    float* maxvals;
    float* weight_out;
    double* partial;
    int accumulate;
    double* result;
    Workspace* ws;
    double* cta_vals;         // [2 * gridDim.x] per-block mse / kl partial sums (workspace tail)
};

// centre of the generated target (uda/dataset/util.py:36-46); multiply-by-reciprocal is bit-identical
// to the division when the stride is a power of two
__device__ __forceinline__ Centre stream_centre(const StreamArgs& a, int map, float& weight) {
    const double jx = a.joints[2 * map], jy = a.joints[2 * map + 1];
    const float vis = a.vis[map];
    double qx, qy;
    if (a.pow2_stride) {
        qx = jx * a.inv_sx;
        qy = jy * a.inv_sy;
    } else {
        qx = __ddiv_rn(jx, a.sx);
        qy = __ddiv_rn(jy, a.sy);
    }
    const double fx = trunc(qx + 0.5), fy = trunc(qy + 0.5);
    const bool inside = (fx >= 0.0) && (fx < static_cast<double>(a.W)) && (fy >= 0.0) && (fy < static_cast<double>(a.H));
    weight = inside ? vis : 0.0f;
    Centre c;
    c.x = inside ? static_cast<int>(fx) : 0;
    c.y = inside ? static_cast<int>(fy) : 0;
    if (!(inside && vis > 0.5f)) c.y = kNoPaste;
    return c;
}

// PCK decision with an fp32 pre-test; the float64 reference arithmetic only near the threshold
__device__ __forceinline__ void stream_pck(const StreamArgs& a, float px, float py, float tx, float ty, int& valid,
                                           int& hit) {
    valid = (tx > 1.0f && ty > 1.0f) ? 1 : 0;
    hit = 0;
    if (!valid) return;
    const float da = (px - tx) * a.inv_nx, db = (py - ty) * a.inv_ny;
    const float d2 = fmaf(da, da, db * db);
    if (d2 < a.thr2_lo) {
        hit = 1;
    } else if (!(d2 > a.thr2_hi)) {
        int v2;
        pck_one(px, py, tx, ty, a.H, a.W, a.thr, v2, hit);
    }
}

template <int NV>
__device__ __forceinline__ void stream_load(const float4* __restrict__ m4, int tile, int lane, float4 (&v)[NV]) {
    const float4* p = m4 + tile * (32 * NV) + lane;
#pragma unroll
    for (int j = 0; j < NV; ++j) v[j] = ldg_stream4(p + j * 32);
}

// running per-lane state of the hot loop
struct StreamAcc {
    float best_v;
    int best_i;
    float m;        // running softmax max (== best_v unless a NaN is present)
    float2 s2;      // sum exp(p - m), two interleaved partial sums
    float2 sp2;     // sum p
    float2 spp2;    // sum p^2
};

template <int NV, int LOSS>
__device__ __forceinline__ void stream_tile(StreamAcc& A, const float4 (&v)[NV], int tile, int lane) {
    // tile maximum (the compiler fuses pairs into FMNMX3)
    float tm = -INFINITY;
#pragma unroll
    for (int j = 0; j < NV; ++j) tm = fmaxf(tm, fmaxf(fmaxf(v[j].x, v[j].y), fmaxf(v[j].z, v[j].w)));
    // first position holding it: scan downwards so the lowest index survives
    int loc = 0;
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

__device__ __forceinline__ float warp_sum(float x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
}

template <int NV, int LOSS>
__global__ void __launch_bounds__(32 * kStreamWarps, 4) pipeline_stream_kernel(const StreamArgs a) {
    extern __shared__ float s_tab[];
    __shared__ double s_cta[2][kStreamWarps];
    __shared__ double s_red[32 * kStreamWarps];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int map = blockIdx.x * kStreamWarps + warp;
    const bool active = map < a.n_maps;
    const float* pm = a.pred + static_cast<size_t>(active ? map : 0) * a.HW;
    const float4* m4 = reinterpret_cast<const float4*>(pm);

    float4 buf0[NV], buf1[NV];
    if (active) stream_load<NV>(m4, 0, lane, buf0);  // in flight during the whole prologue

    load_table(s_tab, a.tab, a.tmp);
    if (lane == 0) {
        s_cta[0][warp] = 0.0;
        s_cta[1][warp] = 0.0;
    }
    __syncthreads();

    if (active) {
        float weight;
        const Centre c = stream_centre(a, map, weight);  // every lane: same inputs, no shuffle needed
        const bool pasted = c.y != kNoPaste;

        // ---- patch pixels of this lane: re-read from the map (L2), target terms in registers -------
        float pt_up = 0.f, pt_ulogu = 0.f, pt_u = 0.f, pt_p = 0.f, pt_e = 0.f;
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
            for (int q = 0; q < kStreamMaxPatch; ++q) {
                if (tv[q] != 0.0f) {
                    const float t = tv[q], p = pv[q], u = t + a.eps;
                    if (LOSS & HP_LOSS_KL) {
                        pt_up = fmaf(u, p, pt_up);
                        pt_ulogu = fmaf(u, lg2_approx(u) * 0.6931471805599453f, pt_ulogu);
                        pt_u += u;
                        pt_p += p;
                    }
                    if (LOSS & HP_LOSS_MSE) pt_e = fmaf(t, t - 2.0f * p, pt_e);  // (p-t)^2 - p^2
                }
            }
        }

        // ---- stream the map: double-buffered register tiles ----------------------------------------
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

        // ---- warp reduction -------------------------------------------------------------------------
        ArgMax am{A.best_v, A.best_i};
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            ArgMax b;
            b.v = __shfl_xor_sync(0xffffffffu, am.v, o);
            b.i = __shfl_xor_sync(0xffffffffu, am.i, o);
            const bool low = (lane & o) == 0;
            am = low ? am_merge(am, b) : am_merge(b, am);
        }
        float s = 0.f;
        if (LOSS & HP_LOSS_KL) {
            const float M = am.v;  // == max of the lane maxima when no NaN is present
            const float ms = (M == -INFINITY) ? 0.0f : M;
            const float scale = (A.m == -INFINITY) ? 0.0f : ex2_approx((A.m - ms) * kLog2e);
            s = warp_sum((A.s2.x + A.s2.y) * scale);
        }
        const float sum_p = warp_sum(A.sp2.x + A.sp2.y);
        if (sum_p != sum_p) {
            // a NaN (or +inf with -inf) is in the map: exact numpy argmax rules, warp-uniform slow path
            ArgMax sx = am_init();
            for (int tile = 0; tile < a.ntiles; ++tile) {
                stream_load<NV>(m4, tile, lane, buf0);
#pragma unroll
                for (int j = 0; j < NV; ++j) am_scan4<true>(sx, buf0[j], tile * (128 * NV) + j * 128 + lane * 4);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                ArgMax b;
                b.v = __shfl_xor_sync(0xffffffffu, sx.v, o);
                b.i = __shfl_xor_sync(0xffffffffu, sx.i, o);
                const bool low = (lane & o) == 0;
                sx = low ? am_merge(sx, b) : am_merge(b, sx);
            }
            am = sx;
            s = __int_as_float(0x7fc00000);  // log_softmax of a map with a NaN is NaN
        }
        float sum_pp = 0.f, r_up = 0.f, r_ulogu = 0.f, r_u = 0.f, r_pp = 0.f, r_e = 0.f;
        if (LOSS & HP_LOSS_MSE) {
            sum_pp = warp_sum(A.spp2.x + A.spp2.y);
            r_e = warp_sum(pt_e);
        }
        if (LOSS & HP_LOSS_KL) {
            r_up = warp_sum(pt_up);
            r_ulogu = warp_sum(pt_ulogu);
            r_u = warp_sum(pt_u);
            r_pp = warp_sum(pt_p);
        }

        // ---- per-map scalars (all lanes compute the same values; lane 0 publishes) ----------------------
        uint32_t qy, qx;
        a.wdiv.divmod(static_cast<uint32_t>(am.i), qy, qx);
        const float keep = (am.v > 0.0f) ? 1.0f : 0.0f;  // NaN -> 0 (keypoint_detection.py:31-34)
        const float px = static_cast<float>(qx) * keep, py = static_cast<float>(qy) * keep;
        const float tx = pasted ? static_cast<float>(c.x) : 0.0f, ty = pasted ? static_cast<float>(c.y) : 0.0f;
        int valid, hit;
        stream_pck(a, px, py, tx, ty, valid, hit);
        double mse = 0.0, kl = 0.0;
        if (LOSS & HP_LOSS_MSE)
            mse = 0.5 * static_cast<double>(weight) * (static_cast<double>(sum_pp) + static_cast<double>(r_e)) /
                  static_cast<double>(a.HW);
        if (LOSS & HP_LOSS_KL) {
            const int n_patch_in = pasted ? (min(c.x + a.tmp, a.W - 1) - max(c.x - a.tmp, 0) + 1) *
                                                (min(c.y + a.tmp, a.H - 1) - max(c.y - a.tmp, 0) + 1)
                                          : 0;
            const float n_bg = static_cast<float>(a.HW - n_patch_in);
            const float Su = fmaf(a.eps, n_bg, r_u);
            const double Sup = static_cast<double>(r_up) +
                               static_cast<double>(a.eps) * (static_cast<double>(sum_p) - static_cast<double>(r_pp));
            const double Sulogu = static_cast<double>(r_ulogu) + static_cast<double>(n_bg) * static_cast<double>(a.eps_log_eps);
            const double lse = static_cast<double>(am.v) + static_cast<double>(logf(s));
            const double L = (Sulogu - Sup) / static_cast<double>(Su) - static_cast<double>(logf(Su)) + lse;
            kl = L * static_cast<double>(weight);
        }
        if (lane == 0) {
            a.pred_xy[2 * map + 0] = px;
            a.pred_xy[2 * map + 1] = py;
            if (a.maxvals) a.maxvals[map] = am.v;
            if (a.weight_out) a.weight_out[map] = weight;
            const int k = map % a.K;
            if (valid) atomicAdd(&a.ws->counts[a.K + k], 1);
            if (hit) atomicAdd(&a.ws->counts[k], 1);
            s_cta[0][warp] = mse;
            s_cta[1][warp] = kl;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {  // fixed warp order: the block's partial sums are deterministic
        double m = 0.0, k = 0.0;
#pragma unroll
        for (int w = 0; w < kStreamWarps; ++w) {
            m += s_cta[0][w];
            k += s_cta[1][w];
        }
        a.cta_vals[blockIdx.x] = m;
        a.cta_vals[gridDim.x + blockIdx.x] = k;
    }

    if (last_block_arrives(&a.ws->counter, gridDim.x)) {
        const volatile double* cv = a.cta_vals;
        const int nb = static_cast<int>(gridDim.x);
        double acc_m = 0.0, acc_k = 0.0;
        for (int i = threadIdx.x; i < nb; i += 32 * kStreamWarps) {
            acc_m += cv[i];
            acc_k += cv[nb + i];
        }
        s_red[threadIdx.x] = acc_m;
        __syncthreads();
        for (int o = 16 * kStreamWarps; o > 0; o >>= 1) {
            if (threadIdx.x < o) s_red[threadIdx.x] += s_red[threadIdx.x + o];
            __syncthreads();
        }
        const double mse_sum = s_red[0];
        __syncthreads();
        s_red[threadIdx.x] = acc_k;
        __syncthreads();
        for (int o = 16 * kStreamWarps; o > 0; o >>= 1) {
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
            P[3] = (a.accumulate ? P[3] : 0.0) + static_cast<double>(a.n_maps) * static_cast<double>(a.HW);
            for (int k = 0; k < 2 * a.K; ++k) {
                P[4 + k] = (a.accumulate ? P[4 + k] : 0.0) + static_cast<double>(cnt[k]);
                cnt[k] = 0;
            }
            if (a.result) pipeline_result_from_partial(P, a.K, a.result);
            a.ws->counter = 0;
        }
    }
}

}  // namespace hp
