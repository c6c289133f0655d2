// hp_loss_variants.cu - SURVEY.md row f3: the two loss weightings next to JointsMSELoss / JointsKLLoss
// (uda/model/loss.py:68-112 JointsMSELoss0, :160-216 JointsKLLoss5; imported at train1.py:24, never called by the drivers).
//
//   JointsMSELoss0: p = (out + 1e-7) / sum(out + 1e-7), t likewise per map, loss = 0.5 * w * (p - t)^2;
//                   'mean' over all elements, 'none' -> mean over HW -> [B,K].
//   JointsKLLoss5:  s[b,k] = w3 / max(w3), w3[b,k] = sum_hw (out / max(out)) * (tgt / max(tgt))  (detached; global maxima),
//                   then the KL loss of (s * out) against (s * tgt), target weights IGNORED;
//                   'mean' over B*K, 'none' -> mean over K -> [B].
// In round 1 both were ATen elementwise chains around the CUDA losses.  Here each is a block-per-map kernel that keeps the
// normalisation in registers (the maps are re-read from L1/L2 between the two or three passes; these variants are not on
// any benchmarked path) with its own backward: autograd never sees an intermediate tensor.
#include "hp_common.cuh"

namespace hp {

constexpr int kLVThreads = 128;

// block-wide sum / max of one float per thread (all threads get the result)
__device__ __forceinline__ float lv_block_sum(float v, float* s_red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.0f;
    for (int w = 0; w < kLVThreads / 32; ++w) t += s_red[w];
    return t;
}
__device__ __forceinline__ float lv_max_nan(float a, float b) {  // torch.max: NaN propagates
    return (a != a || b != b) ? __int_as_float(0x7fc00000) : fmaxf(a, b);
}
__device__ __forceinline__ float lv_block_max(float v, float* s_red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = lv_max_nan(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = s_red[0];
    for (int w = 1; w < kLVThreads / 32; ++w) t = lv_max_nan(t, s_red[w]);
    return t;
}

// ---- JointsMSELoss0 -------------------------------------------------------------------------------------------------
// forward: per_map[m] = 0.5 * w * sum((p - t)^2) / HW.   backward: with g_j = c (p_j - t_j), c = w * go / D,
// d/d out_i = (g_i - sum_j g_j p_j) / S_p  (p_j = (out_j + e) / S_p;  the target has no gradient).
template <bool BWD>
__global__ void __launch_bounds__(kLVThreads) mse0_kernel(const float* __restrict__ out, const float* __restrict__ tgt,
                                                          const float* __restrict__ weight, int n_maps, int K, int HW,
                                                          float* __restrict__ per_map, float* __restrict__ mean, Workspace* ws,
                                                          const float* __restrict__ grad_out, int grad_kind,
                                                          float* __restrict__ grad_in) {
    __shared__ float s_red[kLVThreads / 32];
    const int map = blockIdx.x, t = threadIdx.x;
    const float* po = out + static_cast<size_t>(map) * HW;
    const float* pt = tgt + static_cast<size_t>(map) * HW;
    const float e = 1e-7f;
    float sp = 0.f, st = 0.f;
    for (int i = t; i < HW; i += kLVThreads) {
        sp += __fadd_rn(po[i], e);
        st += __fadd_rn(pt[i], e);
    }
    const float Sp = lv_block_sum(sp, s_red), St = lv_block_sum(st, s_red);
    const float w = weight ? weight[map] : 1.0f;
    if (!BWD) {
        float sd = 0.f;
        for (int i = t; i < HW; i += kLVThreads) {
            const float d = __fsub_rn(__fdiv_rn(__fadd_rn(po[i], e), Sp), __fdiv_rn(__fadd_rn(pt[i], e), St));
            sd = fmaf(d, d, sd);
        }
        const float Sd = lv_block_sum(sd, s_red);
        if (t == 0) {
            const float L = 0.5f * w * Sd / static_cast<float>(HW);
            per_map[map] = L;
            if (mean) fx_acc_add(ws->acc, L);
        }
        if (mean && last_block_arrives_writers(&ws->counter, gridDim.x, t == 0)) {
            if (t == 0) {
                *mean = fx_mean_from_workspace(ws->acc, n_maps);
                ws->counter = 0;
            }
        }
    } else {
        const float go = grad_kind == HP_GRAD_SCALAR ? grad_out[0] : grad_out[map];
        const float D = grad_kind == HP_GRAD_SCALAR ? static_cast<float>(n_maps) * static_cast<float>(HW) : static_cast<float>(HW);
        const float c = w * go / D;
        float sg = 0.f;
        for (int i = t; i < HW; i += kLVThreads) {
            const float p = __fdiv_rn(__fadd_rn(po[i], e), Sp);
            const float g = c * __fsub_rn(p, __fdiv_rn(__fadd_rn(pt[i], e), St));
            sg = fmaf(g, p, sg);
        }
        const float G = lv_block_sum(sg, s_red);
        float* gi = grad_in + static_cast<size_t>(map) * HW;
        for (int i = t; i < HW; i += kLVThreads) {
            const float p = __fdiv_rn(__fadd_rn(po[i], e), Sp);
            const float g = c * __fsub_rn(p, __fdiv_rn(__fadd_rn(pt[i], e), St));
            gi[i] = (g - G) / Sp;
        }
    }
    (void)K;
}

// ---- JointsKLLoss5 --------------------------------------------------------------------------------------------------
// (1) per map: max(out), max(tgt), dot(out, tgt)
__global__ void __launch_bounds__(kLVThreads) kl5_stats_kernel(const float* __restrict__ out, const float* __restrict__ tgt, int HW,
                                                               float* __restrict__ st3) {
    __shared__ float s_red[kLVThreads / 32];
    const int map = blockIdx.x, t = threadIdx.x;
    const float* po = out + static_cast<size_t>(map) * HW;
    const float* pt = tgt + static_cast<size_t>(map) * HW;
    float mo = -INFINITY, mt = -INFINITY, dot = 0.f;
    for (int i = t; i < HW; i += kLVThreads) {
        const float a = po[i], b = pt[i];
        mo = lv_max_nan(mo, a);
        mt = lv_max_nan(mt, b);
        dot = fmaf(a, b, dot);
    }
    mo = lv_block_max(mo, s_red);
    mt = lv_block_max(mt, s_red);
    dot = lv_block_sum(dot, s_red);
    if (t == 0) {
        st3[3 * map + 0] = mo;
        st3[3 * map + 1] = mt;
        st3[3 * map + 2] = dot;
    }
}
// (2) one block: global maxima, w3 = dot / (max_out * max_tgt), scale = w3 / max(w3)      (loss.py:190-196)
__global__ void __launch_bounds__(kLVThreads) kl5_scale_kernel(const float* __restrict__ st3, int n_maps, float* __restrict__ scale) {
    __shared__ float s_red[kLVThreads / 32];
    const int t = threadIdx.x;
    float mo = -INFINITY, mt = -INFINITY;
    for (int i = t; i < n_maps; i += kLVThreads) {
        mo = lv_max_nan(mo, st3[3 * i + 0]);
        mt = lv_max_nan(mt, st3[3 * i + 1]);
    }
    mo = lv_block_max(mo, s_red);
    mt = lv_block_max(mt, s_red);
    float mw = -INFINITY;
    for (int i = t; i < n_maps; i += kLVThreads) {
        const float w3 = __fdiv_rn(__fdiv_rn(st3[3 * i + 2], mo), mt);
        scale[i] = w3;
        mw = lv_max_nan(mw, w3);
    }
    mw = lv_block_max(mw, s_red);
    for (int i = t; i < n_maps; i += kLVThreads) scale[i] = __fdiv_rn(scale[i], mw);
}
// (3) KL of z = s * out against q = (s * tgt + eps) / sum(s * tgt + eps), per map; stats = {lse(z), sum(s tgt + eps)}
template <bool BWD>
__global__ void __launch_bounds__(kLVThreads) kl5_kernel(const float* __restrict__ out, const float* __restrict__ tgt,
                                                         const float* __restrict__ scale, float eps, int n_maps, int K, int HW,
                                                         float* __restrict__ per_map, float* __restrict__ per_sample,
                                                         float* __restrict__ mean, float* __restrict__ stats, Workspace* ws,
                                                         const float* __restrict__ grad_out, int grad_kind,
                                                         float* __restrict__ grad_in) {
    __shared__ float s_red[kLVThreads / 32];
    const int map = blockIdx.x, t = threadIdx.x;
    const float* po = out + static_cast<size_t>(map) * HW;
    const float* pt = tgt + static_cast<size_t>(map) * HW;
    const float s = scale[map];
    if (!BWD) {
        float m = -INFINITY, su = 0.f;
        for (int i = t; i < HW; i += kLVThreads) {
            m = lv_max_nan(m, __fmul_rn(po[i], s));
            su += __fadd_rn(__fmul_rn(pt[i], s), eps);
        }
        const float M = lv_block_max(m, s_red);
        const float Su = lv_block_sum(su, s_red);
        const float Ms = (M == -INFINITY) ? 0.0f : M;
        float se = 0.f, sq = 0.f;  // sum exp(z - M);  sum q (ln q - z)
        for (int i = t; i < HW; i += kLVThreads) {
            const float z = __fmul_rn(po[i], s);
            se += __expf(z - Ms);
            const float q = __fdiv_rn(__fadd_rn(__fmul_rn(pt[i], s), eps), Su);
            // xlogy(q, q) - q * logp: the q ln q term is 0 at q == 0, NaN propagates
            const float qlq = (q == 0.0f) ? 0.0f : q * logf(q);
            sq += qlq - q * z;
        }
        const float Se = lv_block_sum(se, s_red);
        const float Sq = lv_block_sum(sq, s_red);
        // sum_i q_i (ln q_i - (z_i - lse)) = Sq + lse * sum q,  sum q = 1 up to rounding: keep the reference's form
        float sq1 = 0.f;
        for (int i = t; i < HW; i += kLVThreads) sq1 += __fdiv_rn(__fadd_rn(__fmul_rn(pt[i], s), eps), Su);
        const float Q = lv_block_sum(sq1, s_red);
        if (t == 0) {
            const float lse = Ms + logf(Se);
            const float L = fmaf(lse, Q, Sq);
            per_map[map] = L;
            stats[2 * map + 0] = lse;
            stats[2 * map + 1] = Su;
            if (mean) fx_acc_add(ws->acc, L);
        }
        if ((mean || per_sample) && last_block_arrives_writers(&ws->counter, gridDim.x, t == 0)) {
            if (per_sample) per_sample_means(per_map, n_maps / K, K, per_sample, t, kLVThreads);
            if (t == 0) {
                if (mean) *mean = fx_mean_from_workspace(ws->acc, n_maps);
                ws->counter = 0;
            }
        }
    } else {
        // d/d out_i = c * s * (softmax(z)_i * sum q - q_i),  c = go / (B K) ('mean') or go[b] / K ('none')
        const float go = grad_kind == HP_GRAD_SCALAR ? grad_out[0] : grad_out[map / K];
        const float c = go / (grad_kind == HP_GRAD_SCALAR ? static_cast<float>(n_maps) : static_cast<float>(K));
        const float lse = stats[2 * map + 0], Su = stats[2 * map + 1];
        float sq1 = 0.f;
        for (int i = t; i < HW; i += kLVThreads) sq1 += __fdiv_rn(__fadd_rn(__fmul_rn(pt[i], s), eps), Su);
        const float Q = lv_block_sum(sq1, s_red);
        float* gi = grad_in + static_cast<size_t>(map) * HW;
        for (int i = t; i < HW; i += kLVThreads) {
            const float z = __fmul_rn(po[i], s);
            const float q = __fdiv_rn(__fadd_rn(__fmul_rn(pt[i], s), eps), Su);
            gi[i] = c * s * (__expf(z - lse) * Q - q);
        }
    }
}

}  // namespace hp

using namespace hp;

static int lv_check(const char* who, const void* out, const void* tgt, int B, int K, int HW) {
    HP_REQUIRE(out && tgt, HP_ERR_NULL, "%s: null pointer", who);
    HP_REQUIRE(B > 0 && K > 0 && HW > 0 && static_cast<long long>(B) * K < (1ll << 31), HP_ERR_SHAPE, "%s: bad shape B=%d K=%d HW=%d",
               who, B, K, HW);
    return HP_OK;
}

extern "C" HP_API int hp_mse0_fwd(const float* out, const float* tgt, const float* weight, int B, int K, int HW, float* per_map,
                                  float* mean, void* workspace, hp_stream_t stream) {
    if (int rc = lv_check("hp_mse0_fwd", out, tgt, B, K, HW)) return rc;
    HP_REQUIRE(per_map && workspace, HP_ERR_NULL, "hp_mse0_fwd: null pointer");
    mse0_kernel<false><<<B * K, kLVThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        out, tgt, weight, B * K, K, HW, per_map, mean, static_cast<Workspace*>(workspace), nullptr, 0, nullptr);
    return launch_status("hp_mse0_fwd");
}

extern "C" HP_API int hp_mse0_bwd(const float* out, const float* tgt, const float* weight, const float* grad_out, int grad_kind,
                                  int B, int K, int HW, float* grad_in, hp_stream_t stream) {
    if (int rc = lv_check("hp_mse0_bwd", out, tgt, B, K, HW)) return rc;
    HP_REQUIRE(grad_out && grad_in, HP_ERR_NULL, "hp_mse0_bwd: null pointer");
    HP_REQUIRE(grad_kind == HP_GRAD_SCALAR || grad_kind == HP_GRAD_PER_MAP, HP_ERR_ARG, "hp_mse0_bwd: grad_kind %d", grad_kind);
    mse0_kernel<true><<<B * K, kLVThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        out, tgt, weight, B * K, K, HW, nullptr, nullptr, nullptr, grad_out, grad_kind, grad_in);
    return launch_status("hp_mse0_bwd");
}

/* scratch: float [4 * B*K] (per-map max/max/dot, then the per-map scale at scratch + 3*B*K, which the backward re-uses) */
extern "C" HP_API int hp_kl5_fwd(const float* out, const float* tgt, float epsilon, int B, int K, int HW, float* scratch,
                                 float* per_map, float* per_sample, float* mean, float* stats, void* workspace,
                                 hp_stream_t stream) {
    if (int rc = lv_check("hp_kl5_fwd", out, tgt, B, K, HW)) return rc;
    HP_REQUIRE(scratch && per_map && stats && workspace, HP_ERR_NULL, "hp_kl5_fwd: null pointer");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int n = B * K;
    float* scale = scratch + 3 * static_cast<size_t>(n);
    kl5_stats_kernel<<<n, kLVThreads, 0, s>>>(out, tgt, HW, scratch);
    kl5_scale_kernel<<<1, kLVThreads, 0, s>>>(scratch, n, scale);
    kl5_kernel<false><<<n, kLVThreads, 0, s>>>(out, tgt, scale, epsilon, n, K, HW, per_map, per_sample, mean, stats,
                                               static_cast<Workspace*>(workspace), nullptr, 0, nullptr);
    return launch_status("hp_kl5_fwd");
}

extern "C" HP_API int hp_kl5_bwd(const float* out, const float* tgt, float epsilon, const float* scale, const float* stats,
                                 const float* grad_out, int grad_kind, int B, int K, int HW, float* grad_in,
                                 hp_stream_t stream) {
    if (int rc = lv_check("hp_kl5_bwd", out, tgt, B, K, HW)) return rc;
    HP_REQUIRE(scale && stats && grad_out && grad_in, HP_ERR_NULL, "hp_kl5_bwd: null pointer");
    HP_REQUIRE(grad_kind == HP_GRAD_SCALAR || grad_kind == HP_GRAD_PER_SAMPLE, HP_ERR_ARG, "hp_kl5_bwd: grad_kind %d", grad_kind);
    kl5_kernel<true><<<B * K, kLVThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        out, tgt, scale, epsilon, B * K, K, HW, nullptr, nullptr, nullptr, const_cast<float*>(stats), nullptr, grad_out, grad_kind,
        grad_in);
    return launch_status("hp_kl5_bwd");
}
