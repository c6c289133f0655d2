// hp_pipeline_common.cuh - arguments, exact loss accumulation and per-map scalar math shared by the
// three shapes of the fused gen+loss+decode+PCK kernel (generic / stream / coop).
//
// Loss sums are accumulated in 64-bit FIXED POINT (scale 2^40): integer adds are associative, so the
// result is bit-identical for any block schedule, any number of slabs and any sharding over GPUs
// (the all-reduced partial vector is all int64), and no ordered reduction tail is needed.  A per-map
// loss is a float32-accurate value of order 1e-4 .. 1e1; quantising it to 2^-40 ~ 9e-13 is far below
// its own rounding error.  Non-finite per-map losses are counted separately and re-applied at the end.
#pragma once
#include "hp_common.cuh"

namespace hp {

constexpr int kFxShift = 40;
constexpr double kFxLimit = 2097152.0;  // |per-map loss| >= 2^21 is treated as infinite
// partial (int64) = { mse_fx, kl_fx, n_maps, n_elems, hits[K], valid[K],
//                     mse_nan, mse_pinf, mse_ninf, kl_nan, kl_pinf, kl_ninf }


struct PipeArgs {
    const float* pred;
    const double* joints;
    const float* vis;
    int n_maps, K, H, W, HW, ntiles;
    FastDiv wdiv, sdiv;  // by W ; by the patch side (2*tmp+1)
    double sx, sy, inv_sx, inv_sy;
    int pow2_stride;     // joint / stride == joint * inv_stride bit-exactly
    int tmp;
    const float* tab;
    float eps, eps_log_eps;  // eps*ln(eps), 0 when eps == 0
    double thr;
    float thr2_lo, thr2_hi, inv_nx, inv_ny;  // fp32 pre-test of the PCK distance
    int loss_mask;
    float* pred_xy;
    float* maxvals;
    float* weight_out;
    long long* partial;
    int accumulate;
    double* result;
    Workspace* ws;
};

__device__ __forceinline__ float warp_sum(float x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
}
__device__ __forceinline__ float warp_max(float x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x = fmaxf(x, __shfl_xor_sync(0xffffffffu, x, o));
    return x;
}
__device__ __forceinline__ int warp_min_int(int x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x = min(x, __shfl_xor_sync(0xffffffffu, x, o));
    return x;
}
__device__ __forceinline__ ArgMax warp_argmax(ArgMax am, int lane) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        ArgMax b;
        b.v = __shfl_xor_sync(0xffffffffu, am.v, o);
        b.i = __shfl_xor_sync(0xffffffffu, am.i, o);
        am = ((lane & o) == 0) ? am_merge(am, b) : am_merge(b, am);
    }
    return am;
}

// ---- exact accumulation ---------------------------------------------------------------------------
// block-local accumulator (shared memory): [0] mse_fx [1] kl_fx [2..8) non-finite counters
struct BlockLoss {
    long long fx[2];
    int cls[6];
};
__device__ __forceinline__ void block_loss_zero(BlockLoss* b) {
    b->fx[0] = b->fx[1] = 0;
    for (int i = 0; i < 6; ++i) b->cls[i] = 0;
}
// called by ONE thread at a time per block (the finisher of a map)
__device__ __forceinline__ void block_loss_add(BlockLoss* b, int which, double v) {
    if (v != v) b->cls[3 * which + 0] += 1;
    else if (v >= kFxLimit) b->cls[3 * which + 1] += 1;
    else if (v <= -kFxLimit) b->cls[3 * which + 2] += 1;
    else b->fx[which] += __double2ll_rn(ldexp(v, kFxShift));
}
// one thread per block, once: fold the block's sums into the workspace with 64-bit integer atomics
__device__ __forceinline__ void block_loss_flush(const BlockLoss* b, Workspace* ws) {
    for (int w = 0; w < 2; ++w)
        if (b->fx[w] != 0) atomicAdd(&ws->acc[w], static_cast<unsigned long long>(b->fx[w]));
    for (int i = 0; i < 6; ++i)
        if (b->cls[i] != 0) atomicAdd(&ws->acc[2 + i], static_cast<unsigned long long>(b->cls[i]));
}

__device__ __forceinline__ double loss_from_fx(long long fx, long long n_nan, long long n_pinf, long long n_ninf,
                                               long long n_maps) {
    if (n_nan != 0 || (n_pinf != 0 && n_ninf != 0)) return __longlong_as_double(0x7ff8000000000000ll);
    if (n_pinf != 0) return __longlong_as_double(0x7ff0000000000000ll);
    if (n_ninf != 0) return __longlong_as_double(0xfff0000000000000ll);
    return ldexp(static_cast<double>(fx), -kFxShift) / static_cast<double>(n_maps);
}

// result = { mse, kl, avg_acc, cnt, acc[K] }   (JointsMSELoss/JointsKLLoss 'mean', accuracy())
__device__ __forceinline__ void pipeline_result_from_partial(const long long* p, int K, double* result) {
    const long long* cls = p + 4 + 2 * K;
    result[0] = loss_from_fx(p[0], cls[0], cls[1], cls[2], p[2]);
    result[1] = loss_from_fx(p[1], cls[3], cls[4], cls[5], p[2]);
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

// last block, ALL threads: workspace -> partial (= or +=), workspace back to zero, optional finalise.
// One thread per slot so the ~50 counters cost one L2 round trip instead of fifty dependent ones
// (the single-thread version of this was a constant ~20 us tail on every launch, see profiles/).
__device__ __forceinline__ void pipeline_publish(const PipeArgs& a) {
    __shared__ long long s_p[4 + 2 * HP_MAX_K + 6];
    __shared__ double s_acc[HP_MAX_K];
    const int K = a.K, n = 4 + 2 * K + 6;
    const bool add = a.accumulate != 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        long long v;
        if (i < 2) {
            v = static_cast<long long>(*reinterpret_cast<volatile unsigned long long*>(&a.ws->acc[i]));
            a.ws->acc[i] = 0;
        } else if (i == 2) {
            v = a.n_maps;
        } else if (i == 3) {
            v = static_cast<long long>(a.n_maps) * a.HW;
        } else if (i < 4 + 2 * K) {
            v = *reinterpret_cast<volatile int*>(&a.ws->counts[i - 4]);
            a.ws->counts[i - 4] = 0;
        } else {
            const int j = 2 + (i - 4 - 2 * K);
            v = static_cast<long long>(*reinterpret_cast<volatile unsigned long long*>(&a.ws->acc[j]));
            a.ws->acc[j] = 0;
        }
        if (add) v += a.partial[i];
        a.partial[i] = v;
        s_p[i] = v;
    }
    __syncthreads();
    if (a.result) {
        // acc[k] = hits/valid or -1 in parallel; the ordered average over joints serially from shared memory
        for (int k = threadIdx.x; k < K; k += blockDim.x) {
            const long long h = s_p[4 + k], v = s_p[4 + K + k];
            const double acc = v > 0 ? __ddiv_rn(static_cast<double>(h) * 1.0, static_cast<double>(v)) : -1.0;
            s_acc[k] = acc;
            a.result[4 + k] = acc;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            double total = 0.0;
            int cnt = 0;
            for (int k = 0; k < K; ++k)
                if (s_acc[k] >= 0.0) {
                    total = __dadd_rn(total, s_acc[k]);
                    ++cnt;
                }
            const long long* cls = s_p + 4 + 2 * K;
            a.result[0] = loss_from_fx(s_p[0], cls[0], cls[1], cls[2], s_p[2]);
            a.result[1] = loss_from_fx(s_p[1], cls[3], cls[4], cls[5], s_p[2]);
            a.result[2] = cnt != 0 ? __ddiv_rn(total, static_cast<double>(cnt)) : 0.0;
            a.result[3] = static_cast<double>(cnt);
        }
    }
    if (threadIdx.x == 0) a.ws->counter = 0;
}

// "last block done": only thread 0 fences (it is the only writer of block-level results)
__device__ __forceinline__ bool pipeline_last_block(Workspace* ws) {
    __shared__ bool s_last;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        s_last = (atomicAdd(&ws->counter, 1u) == gridDim.x - 1);
        if (s_last) __threadfence();
    }
    __syncthreads();
    return s_last;
}

// ---- per-map scalar math -----------------------------------------------------------------------------
// centre of the generated target (uda/dataset/util.py:36-46); multiplying by the reciprocal is
// bit-identical to the division when the stride is a power of two
// kept out of line so the compiler cannot if-convert the (rare) slow paths into the common path
static __device__ __noinline__ void pipe_centre_divide(double jx, double jy, double sx, double sy, double& qx, double& qy) {
    qx = __ddiv_rn(jx, sx);
    qy = __ddiv_rn(jy, sy);
}
static __device__ __noinline__ int pipe_pck_exact(float px, float py, float tx, float ty, int H, int W, double thr) {
    int valid, hit;
    pck_one(px, py, tx, ty, H, W, thr, valid, hit);
    return hit;
}

__device__ __forceinline__ Centre pipe_centre(const PipeArgs& a, double jx, double jy, float vis, float& weight) {
    double qx, qy;
    if (a.pow2_stride) {
        qx = jx * a.inv_sx;
        qy = jy * a.inv_sy;
    } else {
        pipe_centre_divide(jx, jy, a.sx, a.sy, qx, qy);
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

// PCK decision with an fp32 pre-test; the reference's float64 arithmetic only near the threshold
__device__ __forceinline__ void pipe_pck(const PipeArgs& a, float px, float py, float tx, float ty, int& valid,
                                         int& hit) {
    valid = (tx > 1.0f && ty > 1.0f) ? 1 : 0;
    hit = 0;
    if (!valid) return;
    const float da = (px - tx) * a.inv_nx, db = (py - ty) * a.inv_ny;
    const float d2 = fmaf(da, da, db * db);
    if (d2 < a.thr2_lo) {
        hit = 1;
    } else if (!(d2 > a.thr2_hi)) {
        hit = pipe_pck_exact(px, py, tx, ty, a.H, a.W, a.thr);
    }
}

__device__ __forceinline__ int pipe_patch_area(const PipeArgs& a, Centre c) {
    if (c.y == kNoPaste) return 0;
    return (min(c.x + a.tmp, a.W - 1) - max(c.x - a.tmp, 0) + 1) * (min(c.y + a.tmp, a.H - 1) - max(c.y - a.tmp, 0) + 1);
}

// patch-local target terms of one pixel (u = t + eps):
//   kl: up += u*p, ulogu += u*ln(u), u += u, p += p      mse: e += t*(t - 2p)  ( = (p-t)^2 - p^2 )
struct PatchSums {
    float up, ulogu, u, p, e;
};
template <int LOSS>
__device__ __forceinline__ void patch_pixel(PatchSums& s, float t, float p, float eps) {
    if (t != 0.0f) {
        if (LOSS & HP_LOSS_KL) {
            const float u = t + eps;
            s.up = fmaf(u, p, s.up);
            s.ulogu = fmaf(u, lg2_approx(u) * kLn2, s.ulogu);
            s.u += u;
            s.p += p;
        }
        if (LOSS & HP_LOSS_MSE) s.e = fmaf(t, t - 2.0f * p, s.e);
    }
}

// per-map losses from the reduced sums (all fp32 inputs; closure in float64)
//   sum_exp is relative to vmax;  background pixels have u == eps exactly
template <int LOSS>
__device__ __forceinline__ void pipe_losses(const PipeArgs& a, Centre c, float weight, float vmax, float sum_exp,
                                            float sum_p, float sum_pp, const PatchSums& ps, double& mse, double& kl) {
    mse = 0.0;
    kl = 0.0;
    if (LOSS & HP_LOSS_MSE)  // mean over HW of 0.5*w*(p-t)^2 (loss.py:59-65); sum (p-t)^2 = sum p^2 + sum_patch t(t-2p)
        mse = 0.5 * static_cast<double>(weight) * (static_cast<double>(sum_pp) + static_cast<double>(ps.e)) /
              static_cast<double>(a.HW);
    if (LOSS & HP_LOSS_KL) {
        const float n_bg = static_cast<float>(a.HW - pipe_patch_area(a, c));
        const float Su = fmaf(a.eps, n_bg, ps.u);
        const double Sup = static_cast<double>(ps.up) +
                           static_cast<double>(a.eps) * (static_cast<double>(sum_p) - static_cast<double>(ps.p));
        const double Sulogu = static_cast<double>(ps.ulogu) + static_cast<double>(n_bg) * static_cast<double>(a.eps_log_eps);
        const double lse = static_cast<double>(vmax) + static_cast<double>(lg2_approx(sum_exp) * kLn2);
        // Su == 0 (eps 0 and nothing pasted) -> 0/0 = NaN, like the reference
        const double L = (Sulogu - Sup) / static_cast<double>(Su) - static_cast<double>(lg2_approx(Su) * kLn2) + lse;
        kl = L * static_cast<double>(weight);
    }
}

}  // namespace hp
