// hp_variants.cu - SURVEY.md row f3: the label-fusing disparity variants RegressionDisparity2/3/5/6/7/8
// (uda/model/regda_4.py:145-645; imported at train1.py:19, never called by the drivers).
//
// They all build, per SAMPLE, one map `label_p` from the summed Gaussian pseudo-labels of up to three predictions
// (y, label_1, label_2) and then the per-joint targets
//     gt[b,k] = Gaussian at argmax(y[b,k])              gf[b,k] = clip(label_p[b] - 10 gt[b,k], 0, 1)
// which go to the criterion like any other target map.  In round 1 this was a chain of ~15 elementwise / reduction ATen
// launches over [B,K,H,W] temporaries; here it is ONE kernel after the three decodes: a block per sample rebuilds the three
// per-sample sums from the 3K centres in registers (ascending joint order, like torch's sum over dim 1), applies the
// variant's rule, takes the per-sample maximum where the rule normalises by it, and writes gt and gf (the criterion is
// arbitrary, so the maps are materialised - 2 * K * H*W*4 bytes per sample written, nothing read but 3K centres).
//   rd2: lp = (s1 + s0 + s2) / max            rd3: lp = (c(s0) + c(s1) + c(s2)) / max        c = clip to [0,1]
//   rd5: lp = c(p1 + c(p2-p1) + c(p3-p1))     rd6: lp = c(p1 + c(p2-p1))                      p_i = c(s_i)
//   rd7: lp = (c(s1) + c(s0)) / max           rd8: lp = c(p1 + c(sum_k c(G1_k - G0_k)) + c(sum_k c(G2_k - G0_k)))
#include "hp_common.cuh"

namespace hp {

__device__ __forceinline__ float clip01v(float x) {  // torch.clip(min=0, max=1): NaN propagates
    float y;
    asm("max.NaN.f32 %0, %1, 0f00000000;\n\tmin.NaN.f32 %0, %0, 0f3F800000;" : "=f"(y) : "f"(x));
    return y;
}

constexpr int kLFThreads = 256;
constexpr int kLFMaxPix = 16;  // pixels per thread: H*W <= 4096

__global__ void __launch_bounds__(kLFThreads) label_fusion_kernel(const int32_t* __restrict__ c0, const int32_t* __restrict__ c1,
                                                                  const int32_t* __restrict__ c2, int rule, int K, int H, int W,
                                                                  int tmp, const float* __restrict__ tab,
                                                                  float* __restrict__ gt, float* __restrict__ gf,
                                                                  float* __restrict__ label_p) {
    extern __shared__ float s_lf[];
    __shared__ Centre s_c[3][HP_MAX_K];
    __shared__ float s_max[kLFThreads / 32];
    float* s_tab = s_lf;
    const int b = blockIdx.x, t = threadIdx.x, HW = H * W;
    load_table(s_tab, tab, tmp);
    for (int i = t; i < 3 * K; i += kLFThreads) {
        const int src = i / K, k = i - src * K;
        const int32_t* c = src == 0 ? c0 : (src == 1 ? c1 : c2);
        Centre cc{0, kNoPaste};
        if (c) cc = Centre{c[2 * (b * K + k)], c[2 * (b * K + k) + 1]};
        s_c[src][k] = cc;
    }
    __syncthreads();
    const bool three = c2 != nullptr;
    float lp[kLFMaxPix];
    float lmax = -INFINITY;
#pragma unroll
    for (int j = 0; j < kLFMaxPix; ++j) {
        const int e = t + j * kLFThreads;
        lp[j] = 0.0f;
        if (e >= HW) continue;
        const int y = e / W, x = e - y * W;
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, d1 = 0.f, d2 = 0.f;
        for (int k = 0; k < K; ++k) {  // ascending joint order = torch.sum(dim=1)
            const float g0 = patch_at(s_tab, tmp, s_c[0][k], x, y);
            const float g1 = patch_at(s_tab, tmp, s_c[1][k], x, y);
            const float g2 = three ? patch_at(s_tab, tmp, s_c[2][k], x, y) : 0.0f;
            s0 += g0; s1 += g1; s2 += g2;
            if (rule == HP_LF_RD8) {
                d1 += clip01v(g1 - g0);
                d2 += clip01v(g2 - g0);
            }
        }
        const float p1 = clip01v(s0), p2 = clip01v(s1), p3 = clip01v(s2);
        float v;
        switch (rule) {
            case HP_LF_RD2: v = (s1 + s0) + s2; break;                                 // regda_4.py:203-206
            case HP_LF_RD3: v = (p1 + p2) + p3; break;                                 // :280-283
            case HP_LF_RD5: v = clip01v((p1 + clip01v(p2 - p1)) + clip01v(p3 - p1)); break;   // :409-416
            case HP_LF_RD6: v = clip01v(p1 + clip01v(p2 - p1)); break;                 // :479-484
            case HP_LF_RD7: v = p2 + p1; break;                                        // :553-555
            default: v = clip01v((p1 + clip01v(d1)) + clip01v(d2)); break;             // rd8 :625-634
        }
        lp[j] = v;
        lmax = fmaxf(lmax, v);
        if (v != v) lmax = v;  // torch.max propagates NaN
    }
    const bool normalise = rule == HP_LF_RD2 || rule == HP_LF_RD3 || rule == HP_LF_RD7;
    float M = 1.0f;
    if (normalise) {
        // per-sample maximum (NaN wins, like torch.max)
        float m = lmax;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float other = __shfl_xor_sync(0xffffffffu, m, o);
            m = (m != m || other != other) ? __int_as_float(0x7fc00000) : fmaxf(m, other);
        }
        if ((t & 31) == 0) s_max[t >> 5] = m;
        __syncthreads();
        M = s_max[0];
        for (int w = 1; w < kLFThreads / 32; ++w) {
            const float o = s_max[w];
            M = (M != M || o != o) ? __int_as_float(0x7fc00000) : fmaxf(M, o);
        }
    }
#pragma unroll
    for (int j = 0; j < kLFMaxPix; ++j) {
        const int e = t + j * kLFThreads;
        if (e >= HW) continue;
        const float v = normalise ? __fdiv_rn(lp[j], M) : lp[j];
        if (label_p) label_p[static_cast<size_t>(b) * HW + e] = v;
        const int y = e / W, x = e - y * W;
        for (int k = 0; k < K; ++k) {
            const float g = patch_at(s_tab, tmp, s_c[0][k], x, y);
            const size_t o = (static_cast<size_t>(b) * K + k) * HW + e;
            if (gt) gt[o] = g;
            if (gf) gf[o] = clip01v(__fsub_rn(v, __fmul_rn(g, 10.0f)));               // regda_4.py:210-213
        }
    }
}

}  // namespace hp

using namespace hp;

extern "C" HP_API int hp_label_fusion(const int32_t* centres_y, const int32_t* centres_1, const int32_t* centres_2,
                                      int rule, int B, int K, int H, int W, int tmp, const float* tab, float* gt,
                                      float* gf, float* label_p, hp_stream_t stream) {
    HP_REQUIRE(centres_y && centres_1 && tab && (gt || gf || label_p), HP_ERR_NULL, "hp_label_fusion: null pointer");
    HP_REQUIRE(rule == HP_LF_RD2 || rule == HP_LF_RD3 || rule == HP_LF_RD5 || rule == HP_LF_RD6 || rule == HP_LF_RD7 ||
                   rule == HP_LF_RD8,
               HP_ERR_ARG, "hp_label_fusion: rule %d", rule);
    const bool needs_two = rule == HP_LF_RD2 || rule == HP_LF_RD3 || rule == HP_LF_RD5 || rule == HP_LF_RD8;
    HP_REQUIRE(!needs_two || centres_2, HP_ERR_NULL, "hp_label_fusion: this rule fuses two extra labels");
    HP_REQUIRE(B > 0 && K > 0 && K <= HP_MAX_K && H > 0 && W > 0 && H * W <= kLFThreads * kLFMaxPix && tmp >= 0 && tmp <= 64,
               HP_ERR_SHAPE, "hp_label_fusion: bad shape B=%d K=%d H=%d W=%d (H*W <= %d)", B, K, H, W, kLFThreads * kLFMaxPix);
    label_fusion_kernel<<<B, kLFThreads, table_bytes(tmp), static_cast<cudaStream_t>(stream)>>>(
        centres_y, centres_1, needs_two ? centres_2 : nullptr, rule, K, H, W, tmp, tab, gt, gf, label_p);
    return launch_status("hp_label_fusion");
}
