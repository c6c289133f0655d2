// hp_loss.cu - JointsMSELoss (a3) and JointsKLLoss (a4), forward and backward.
//
// Replaces uda/model/loss.py:55-65 and :145-158.  The reference runs ~5 / ~8 ATen kernels with
// 3-4 full-size temporaries; here each direction is ONE pass: prediction and target tiles are
// loaded once into registers (128-bit no-allocate loads), every per-map statistic is accumulated
// in that pass (online softmax), and the last block reduces the per-map values in a fixed order.
//
// KL per map, with u_i = t_i + eps, S = sum u_i, q = u / S:
//   L = sum q log q - sum q p + logsumexp(p) = (sum u log u - sum u p) / S - log S + lse
// Roofline: HBM.  Algorithmic bytes per map: forward 2*HW*4 read; backward 2*HW*4 read + HW*4 written.
#include "hp_common.cuh"
#include "hp_dispatch.cuh"
#include "hp_loss_staged.cuh"

namespace hp {

// ---------------------------------------------------------------------------------------------
// forward kernels
// ---------------------------------------------------------------------------------------------
template <int TPM, int NV, int MODE, int MPB, bool IS_KL>
__global__ void __launch_bounds__(TPM* MPB)
    loss_fwd_kernel(const float* __restrict__ output, const float* __restrict__ target, const float* __restrict__ weight,
                    float eps, int n_maps, int K, int HW, float* __restrict__ per_map, float* __restrict__ per_sample,
                    float* __restrict__ mean, float* __restrict__ stats, Workspace* __restrict__ ws) {
    constexpr int NS = 3;
    __shared__ Stats<NS> scratch[TPM > 32 ? TPM / 32 + 1 : 1];
    const int g = threadIdx.x / TPM, t = threadIdx.x % TPM;
    const int map = blockIdx.x * MPB + g;
    if (map < n_maps) {
        const float* pm = output + static_cast<size_t>(map) * HW;
        const float* tm = target + static_cast<size_t>(map) * HW;
        Stats<NS> st;
        stats_init(st);
        const int ntiles = (MODE == WALK_EXACT) ? 1 : tiles_for<TPM, NV>(HW);
        for (int tile = 0; tile < ntiles; ++tile) {
            float4 p[NV], q[NV];
            load_tile<TPM, NV, MODE>(pm, HW, tile, t, IS_KL ? -INFINITY : 0.0f, p);
            load_tile<TPM, NV, MODE>(tm, HW, tile, t, 0.0f, q);
            if (IS_KL) softmax_tile<NV>(st.m, st.s, p);
#pragma unroll
            for (int j = 0; j < NV; ++j) {
                const int idx0 = tile * (TPM * NV * 4) + (j * TPM + t) * 4;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    if (MODE != WALK_EXACT && idx0 + c >= HW) continue;
                    const float pv = f4_get(p[j], c), tv = f4_get(q[j], c);
                    if (IS_KL) {
                        kl_elem(st.sum, pv, tv + eps);
                    } else {
                        const float d = pv - tv;
                        st.sum[0] = fmaf(d, d, st.sum[0]);
                    }
                }
            }
        }
        group_reduce<TPM, NS, false, IS_KL, false>(st, scratch);
        if (t == 0) {
            const float w = weight ? weight[map] : 1.0f;
            if (IS_KL) {
                float lse;
                const double L = kl_finish(st.m, st.s, st.sum[0], st.sum[1], st.sum[2], lse);
                const float Lw = static_cast<float>(L * static_cast<double>(w));
                per_map[map] = Lw;
                if (mean) fx_acc_add(ws->acc, Lw);
                if (stats) {
                    stats[2 * map + 0] = lse;
                    stats[2 * map + 1] = st.sum[0];
                }
            } else {
                // mean over HW of 0.5*w*(p-t)^2  (loss.py:59-65)
                const float Lw = 0.5f * w * (st.sum[0] / static_cast<float>(HW));
                per_map[map] = Lw;
                if (mean) fx_acc_add(ws->acc, Lw);
            }
        }
    }
    if (mean == nullptr && per_sample == nullptr) return;
    if (last_block_arrives_writers(&ws->counter, gridDim.x, t == 0 && map < n_maps)) {
        if (per_sample) per_sample_means(per_map, n_maps / K, K, per_sample, threadIdx.x, TPM * MPB);  // KL 'none'
        if (mean && threadIdx.x == 0) {
            // MSE 'mean' = mean over all elements = mean over maps of the per-map means (equal HW)
            *mean = fx_mean_from_workspace(ws->acc, n_maps);
        }
        if (threadIdx.x == 0) ws->counter = 0;
    }
}

// ---------------------------------------------------------------------------------------------
// backward kernels (elementwise, one map per thread group so the per-map coefficient is uniform)
//   MSE: d/dp = g * w * (p - t) / N          N = B*K*HW ('mean')  or HW ('none', g per map)
//   KL : d/dp = g * w / D * (softmax(p) - q) D = B*K   ('mean')  or K  ('none', g per sample)
// ---------------------------------------------------------------------------------------------
template <int TPM, int NV, int MODE, int MPB, bool IS_KL>
__global__ void __launch_bounds__(TPM* MPB)
    loss_bwd_kernel(const float* __restrict__ output, const float* __restrict__ target, const float* __restrict__ weight,
                    float eps, const float* __restrict__ stats, const float* __restrict__ grad_out, int grad_kind,
                    int n_maps, int K, int HW, float* __restrict__ grad_in) {
    const int g = threadIdx.x / TPM, t = threadIdx.x % TPM;
    const int map = blockIdx.x * MPB + g;
    if (map >= n_maps) return;
    const float w = weight ? weight[map] : 1.0f;
    float go, denom;
    if (grad_kind == HP_GRAD_SCALAR) {
        go = grad_out[0];
        denom = IS_KL ? static_cast<float>(n_maps) : static_cast<float>(n_maps) * static_cast<float>(HW);
    } else if (grad_kind == HP_GRAD_PER_MAP) {
        go = grad_out[map];
        denom = IS_KL ? 1.0f : static_cast<float>(HW);
    } else {
        go = grad_out[map / K];
        denom = IS_KL ? static_cast<float>(K) : static_cast<float>(K) * static_cast<float>(HW);
    }
    const float coef = go * w / denom;
    float lse = 0.f, invS = 0.f;
    if (IS_KL) {
        lse = stats[2 * map + 0];
        invS = 1.0f / stats[2 * map + 1];
    }
    const float lb = -lse * kLog2e;
    const float* pm = output + static_cast<size_t>(map) * HW;
    const float* tm = target + static_cast<size_t>(map) * HW;
    float* gm = grad_in + static_cast<size_t>(map) * HW;
    const int ntiles = (MODE == WALK_EXACT) ? 1 : tiles_for<TPM, NV>(HW);
    for (int tile = 0; tile < ntiles; ++tile) {
        float4 p[NV], q[NV];
        load_tile<TPM, NV, MODE>(pm, HW, tile, t, 0.0f, p);
        load_tile<TPM, NV, MODE>(tm, HW, tile, t, 0.0f, q);
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            const int idx0 = tile * (TPM * NV * 4) + (j * TPM + t) * 4;
            float r[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const float pv = f4_get(p[j], c), tv = f4_get(q[j], c);
                if (IS_KL) r[c] = coef * (exp2f(fmaf(pv, kLog2e, lb)) - (tv + eps) * invS);
                else r[c] = coef * (pv - tv);
            }
            if (MODE == WALK_SCALAR) {
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    if (idx0 + c < HW) gm[idx0 + c] = r[c];
            } else if (MODE == WALK_EXACT || idx0 < HW) {
                stg_stream4(reinterpret_cast<float4*>(gm + idx0), make_float4(r[0], r[1], r[2], r[3]));
            }
        }
    }
}

template <bool IS_KL>
struct LossFwdLaunch {
    const float *output, *target, *weight;
    float eps;
    int n_maps, K, HW;
    float *per_map, *per_sample, *mean, *stats;
    Workspace* ws;
    cudaStream_t stream;
    template <int TPM, int NV, int MODE, int MPB>
    void run() const {
        const int grid = (n_maps + MPB - 1) / MPB;
        loss_fwd_kernel<TPM, NV, MODE, MPB, IS_KL><<<grid, TPM * MPB, 0, stream>>>(
            output, target, weight, eps, n_maps, K, HW, per_map, per_sample, mean, stats, ws);
    }
};

template <bool IS_KL>
struct LossBwdLaunch {
    const float *output, *target, *weight;
    float eps;
    const float *stats, *grad_out;
    int grad_kind, n_maps, K, HW;
    float* grad_in;
    cudaStream_t stream;
    template <int TPM, int NV, int MODE, int MPB>
    void run() const {
        const int grid = (n_maps + MPB - 1) / MPB;
        loss_bwd_kernel<TPM, NV, MODE, MPB, IS_KL><<<grid, TPM * MPB, 0, stream>>>(
            output, target, weight, eps, stats, grad_out, grad_kind, n_maps, K, HW, grad_in);
    }
};

static int check_loss_args(const char* who, const void* output, const void* target, int B, int K, int HW) {
    HP_REQUIRE(output && target, HP_ERR_NULL, "%s: null pointer", who);
    HP_REQUIRE(B > 0 && K > 0 && HW > 0 && HW < (1 << 30) && static_cast<long long>(B) * K < (1ll << 31), HP_ERR_SHAPE,
               "%s: bad shape B=%d K=%d HW=%d", who, B, K, HW);
    HP_REQUIRE(aligned4(output) && aligned4(target), HP_ERR_ALIGN, "%s: misaligned input", who);
    return HP_OK;
}

}  // namespace hp

using namespace hp;

extern "C" HP_API int hp_mse_fwd(const float* output, const float* target, const float* weight, int B, int K, int HW,
                                 float* per_map, float* mean, void* workspace, hp_stream_t stream) {
    if (int rc = check_loss_args("hp_mse_fwd", output, target, B, K, HW)) return rc;
    HP_REQUIRE(per_map && workspace, HP_ERR_NULL, "hp_mse_fwd: null output");
    {   // 64x64 maps: warp-private copy-engine stages (hp_loss_staged.cuh); everything else: block per map
        const LossStagedArgs sa{output, target, weight, 0.f, B * K, K, per_map, nullptr, mean, nullptr, static_cast<Workspace*>(workspace)};
        const int rc = launch_loss_fwd_staged<false>(sa, HW, static_cast<cudaStream_t>(stream), "hp_mse_fwd");
        if (rc != 1) return rc;
    }
    LossFwdLaunch<false> l{output, target, weight, 0.f, B * K, K, HW, per_map, nullptr, mean, nullptr,
                           static_cast<Workspace*>(workspace), static_cast<cudaStream_t>(stream)};
    dispatch_map_walk<true>(HW, aligned16(output) && aligned16(target), l);
    return launch_status("hp_mse_fwd");
}

extern "C" HP_API int hp_mse_bwd(const float* output, const float* target, const float* weight, const float* grad_out,
                                 int grad_kind, int B, int K, int HW, float* grad_in, hp_stream_t stream) {
    if (int rc = check_loss_args("hp_mse_bwd", output, target, B, K, HW)) return rc;
    HP_REQUIRE(grad_out && grad_in, HP_ERR_NULL, "hp_mse_bwd: null gradient pointer");
    HP_REQUIRE(grad_kind == HP_GRAD_SCALAR || grad_kind == HP_GRAD_PER_MAP, HP_ERR_ARG, "hp_mse_bwd: grad_kind %d",
               grad_kind);
    LossBwdLaunch<false> l{output, target, weight, 0.f, nullptr, grad_out, grad_kind, B * K, K, HW, grad_in,
                           static_cast<cudaStream_t>(stream)};
    dispatch_map_walk<true>(HW, aligned16(output) && aligned16(target) && aligned16(grad_in), l);
    return launch_status("hp_mse_bwd");
}

extern "C" HP_API int hp_kl_fwd(const float* output, const float* target, const float* weight, float epsilon, int B,
                                int K, int HW, float* per_map, float* per_sample, float* mean, float* stats,
                                void* workspace, hp_stream_t stream) {
    if (int rc = check_loss_args("hp_kl_fwd", output, target, B, K, HW)) return rc;
    HP_REQUIRE(per_map && workspace, HP_ERR_NULL, "hp_kl_fwd: null output");
    {
        const LossStagedArgs sa{output, target, weight, epsilon, B * K, K, per_map, per_sample, mean, stats, static_cast<Workspace*>(workspace)};
        const int rc = launch_loss_fwd_staged<true>(sa, HW, static_cast<cudaStream_t>(stream), "hp_kl_fwd");
        if (rc != 1) return rc;
    }
    LossFwdLaunch<true> l{output, target, weight, epsilon, B * K, K, HW, per_map, per_sample, mean, stats,
                          static_cast<Workspace*>(workspace), static_cast<cudaStream_t>(stream)};
    dispatch_map_walk<true>(HW, aligned16(output) && aligned16(target), l);
    return launch_status("hp_kl_fwd");
}

extern "C" HP_API int hp_kl_bwd(const float* output, const float* target, const float* weight, float epsilon,
                                const float* stats, const float* grad_out, int grad_kind, int B, int K, int HW,
                                float* grad_in, hp_stream_t stream) {
    if (int rc = check_loss_args("hp_kl_bwd", output, target, B, K, HW)) return rc;
    HP_REQUIRE(stats && grad_out && grad_in, HP_ERR_NULL, "hp_kl_bwd: null pointer");
    HP_REQUIRE(grad_kind == HP_GRAD_SCALAR || grad_kind == HP_GRAD_PER_SAMPLE, HP_ERR_ARG, "hp_kl_bwd: grad_kind %d",
               grad_kind);
    LossBwdLaunch<true> l{output, target, weight, epsilon, stats, grad_out, grad_kind, B * K, K, HW, grad_in,
                          static_cast<cudaStream_t>(stream)};
    dispatch_map_walk<true>(HW, aligned16(output) && aligned16(target) && aligned16(grad_in), l);
    return launch_status("hp_kl_bwd");
}
