// hp_fusion.cu - multiscale heatmap fusion (a12) and the fused fuse+decode+PCK of config 4.
//
// Replaces the inline statements of train1.py:410-424 (== test.py:362-376):
//     target5 = 0.5 * Upsample(64)(y_adv3) + Upsample(64)(y_adv2);  target0 = Upsample(32)(y_adv3)
// nn.Upsample(mode='bilinear') is align_corners=False:  src = max(scale*(dst+0.5)-0.5, 0),
// scale = in/out (fp32), i0 = floor(src), i1 = i0 + (i0 < in-1), l1 = src - i0, l0 = 1 - l1,
// out = l0y*(l0x*v00 + l1x*v01) + l1y*(l0x*v10 + l1x*v11)            (SURVEY.md appendix A8).
// The reference materialises three upsampled tensors and two elementwise temporaries; here the
// 4-tap gathers from the small maps (L1-resident: 1-16 KB per map) and the blend happen in
// registers and the fused map is written once - or, for hp_fuse_decode_pck, never written.
// Roofline: HBM.  Bytes per map: sum of the source maps read [+ H*W*4 written].
#include "hp_common.cuh"
#include "hp_decode.cuh"
#include "hp_dispatch.cuh"

namespace hp {

struct Tap {
    int i0, i1;
    float l0, l1;
};

__device__ __forceinline__ Tap make_tap(int o, float scale, int in_size, int out_size) {
    Tap tp;
    if (in_size == out_size) {
        tp.i0 = tp.i1 = o;
        tp.l0 = 1.0f;
        tp.l1 = 0.0f;
        return tp;
    }
    float src = __fsub_rn(__fmul_rn(scale, static_cast<float>(o) + 0.5f), 0.5f);
    if (src < 0.0f) src = 0.0f;
    int i0 = static_cast<int>(floorf(src));
    if (i0 > in_size - 1) i0 = in_size - 1;
    tp.i0 = i0;
    tp.i1 = i0 + ((i0 < in_size - 1) ? 1 : 0);
    float l1 = src - static_cast<float>(i0);
    l1 = fminf(fmaxf(l1, 0.0f), 1.0f);
    tp.l1 = l1;
    tp.l0 = 1.0f - l1;
    return tp;
}

__device__ __forceinline__ float bilerp(const float* __restrict__ src, int w, Tap ty, Tap tx) {
    const float* r0 = src + ty.i0 * w;
    const float* r1 = src + ty.i1 * w;
    const float top = __fadd_rn(__fmul_rn(tx.l0, __ldg(r0 + tx.i0)), __fmul_rn(tx.l1, __ldg(r0 + tx.i1)));
    const float bot = __fadd_rn(__fmul_rn(tx.l0, __ldg(r1 + tx.i0)), __fmul_rn(tx.l1, __ldg(r1 + tx.i1)));
    return __fadd_rn(__fmul_rn(ty.l0, top), __fmul_rn(ty.l1, bot));
}

struct FuseSrc {
    const float* lo;
    int hl, wl;
    float a_lo, sy_lo, sx_lo;
    const float* mid;
    int hm, wm;
    float a_mid, sy_mid, sx_mid;
    const float* hi;
    float a_hi;
    int H, W;
    FastDiv wdiv;
};

// fused value at (x, y) of map `map`:  a_lo*up(lo) [+ a_mid*up(mid)] [+ a_hi*hi]   (left-to-right fp32)
__device__ __forceinline__ float fused_at(const FuseSrc& f, size_t map, int x, int y, float hi_val) {
    const Tap tyl = make_tap(y, f.sy_lo, f.hl, f.H), txl = make_tap(x, f.sx_lo, f.wl, f.W);
    float r = __fmul_rn(f.a_lo, bilerp(f.lo + map * f.hl * f.wl, f.wl, tyl, txl));
    if (f.mid) {
        const Tap tym = make_tap(y, f.sy_mid, f.hm, f.H), txm = make_tap(x, f.sx_mid, f.wm, f.W);
        r = __fadd_rn(r, __fmul_rn(f.a_mid, bilerp(f.mid + map * f.hm * f.wm, f.wm, tym, txm)));
    }
    if (f.hi) r = __fadd_rn(r, __fmul_rn(f.a_hi, hi_val));
    return r;
}

// ---------------------------------------------------------------------------------------------
// materialising kernel: one thread per 4 consecutive outputs of a row (or per element if W % 4 != 0)
// ---------------------------------------------------------------------------------------------
template <bool VEC>
__global__ void __launch_bounds__(256) fuse_kernel(const FuseSrc f, int n_maps, float* __restrict__ out) {
    const int HW = f.H * f.W;
    const size_t total = static_cast<size_t>(n_maps) * HW / (VEC ? 4 : 1);
    for (size_t v = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; v < total;
         v += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const size_t e = VEC ? v * 4 : v;
        const size_t map = e / HW;
        const uint32_t rem = static_cast<uint32_t>(e - map * HW);
        uint32_t y, x0;
        f.wdiv.divmod(rem, y, x0);
        if (VEC) {
            float4 h = make_float4(0.f, 0.f, 0.f, 0.f);
            if (f.hi) h = ldg_stream4(reinterpret_cast<const float4*>(f.hi + e));
            float4 r;
            r.x = fused_at(f, map, x0 + 0, y, h.x);
            r.y = fused_at(f, map, x0 + 1, y, h.y);
            r.z = fused_at(f, map, x0 + 2, y, h.z);
            r.w = fused_at(f, map, x0 + 3, y, h.w);
            stg_stream4(reinterpret_cast<float4*>(out + e), r);
        } else {
            out[e] = fused_at(f, map, x0, y, f.hi ? f.hi[e] : 0.0f);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// fuse + decode + PCK, fused map kept in registers (BASELINE.json configs[3])
// ---------------------------------------------------------------------------------------------
template <int TPM, int NV, int MODE>
struct FusedTileLoader {
    const FuseSrc& f;
    size_t map;
    int HW, t;
    __device__ __forceinline__ void operator()(int tile, float fill, float4 (&v)[NV]) const {
        float4 h[NV];
        if (f.hi) load_tile<TPM, NV, MODE>(f.hi + map * HW, HW, tile, t, 0.0f, h);
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            const int idx0 = tile * (TPM * NV * 4) + (j * TPM + t) * 4;
            float r[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                r[c] = fill;
                if (idx0 + c < HW) {
                    uint32_t y, x;
                    f.wdiv.divmod(static_cast<uint32_t>(idx0 + c), y, x);
                    r[c] = fused_at(f, map, x, y, f.hi ? f4_get(h[j], c) : 0.0f);
                }
            }
            v[j] = make_float4(r[0], r[1], r[2], r[3]);
        }
    }
};

template <int TPM, int NV, int MODE, int MPB>
__global__ void __launch_bounds__(TPM* MPB)
    fuse_decode_pck_kernel(const FuseSrc f, const float* __restrict__ tgt_xy, int n_maps, int K, double thr,
                           float* __restrict__ pred_xy, float* __restrict__ maxvals, int32_t* __restrict__ counts_out,
                           double* __restrict__ acc_out, Workspace* __restrict__ ws) {
    __shared__ Stats<0> scratch[TPM > 32 ? TPM / 32 + 1 : 1];
    const int HW = f.H * f.W;
    const int g = threadIdx.x / TPM, t = threadIdx.x % TPM;
    const int map = blockIdx.x * MPB + g;
    if (map < n_maps) {
        FusedTileLoader<TPM, NV, MODE> ld{f, static_cast<size_t>(map), HW, t};
        const ArgMax a = decode_tiles<TPM, NV, MODE>(ld, HW, t, scratch);
        if (t == 0) {
            float px, py;
            decode_xy(a, f.W, px, py);
            pred_xy[2 * map + 0] = px;
            pred_xy[2 * map + 1] = py;
            if (maxvals) maxvals[map] = a.v;
            int valid, hit;
            pck_one(px, py, tgt_xy[2 * map], tgt_xy[2 * map + 1], f.H, f.W, thr, valid, hit);
            const int k = map % K;
            if (valid) atomicAdd(&ws->counts[K + k], 1);
            if (hit) atomicAdd(&ws->counts[k], 1);
        }
    }
    if (last_block_arrives(&ws->counter, gridDim.x)) pck_publish(ws, K, counts_out, acc_out);
}

struct FuseDecodeLaunch {
    FuseSrc f;
    const float* tgt_xy;
    int n_maps, K;
    double thr;
    float* pred_xy;
    float* maxvals;
    int32_t* counts;
    double* acc;
    Workspace* ws;
    cudaStream_t stream;
    template <int TPM, int NV, int MODE, int MPB>
    void run() const {
        const int grid = (n_maps + MPB - 1) / MPB;
        fuse_decode_pck_kernel<TPM, NV, MODE, MPB>
            <<<grid, TPM * MPB, 0, stream>>>(f, tgt_xy, n_maps, K, thr, pred_xy, maxvals, counts, acc, ws);
    }
};

static int make_src(const char* who, const float* lo, int hl, int wl, float a_lo, const float* mid, int hm, int wm,
                    float a_mid, const float* hi, float a_hi, int H, int W, FuseSrc& f) {
    HP_REQUIRE(lo, HP_ERR_NULL, "%s: lo is null", who);
    HP_REQUIRE(hl > 0 && wl > 0 && H > 0 && W > 0 && static_cast<long long>(H) * W < (1ll << 30), HP_ERR_SHAPE,
               "%s: bad shape lo=%dx%d out=%dx%d", who, hl, wl, H, W);
    HP_REQUIRE(!mid || (hm > 0 && wm > 0), HP_ERR_SHAPE, "%s: bad mid shape %dx%d", who, hm, wm);
    f.lo = lo; f.hl = hl; f.wl = wl; f.a_lo = a_lo;
    f.sy_lo = static_cast<float>(hl) / static_cast<float>(H);
    f.sx_lo = static_cast<float>(wl) / static_cast<float>(W);
    f.mid = mid; f.hm = hm; f.wm = wm; f.a_mid = a_mid;
    f.sy_mid = mid ? static_cast<float>(hm) / static_cast<float>(H) : 1.0f;
    f.sx_mid = mid ? static_cast<float>(wm) / static_cast<float>(W) : 1.0f;
    f.hi = hi; f.a_hi = a_hi; f.H = H; f.W = W;
    f.wdiv = FastDiv(static_cast<uint32_t>(W));
    return HP_OK;
}

}  // namespace hp

using namespace hp;

extern "C" HP_API int hp_fuse_multiscale(const float* lo, int hl, int wl, float a_lo, const float* mid, int hm, int wm,
                                         float a_mid, const float* hi, float a_hi, int n_maps, int H, int W, float* out,
                                         hp_stream_t stream) {
    FuseSrc f;
    if (int rc = make_src("hp_fuse_multiscale", lo, hl, wl, a_lo, mid, hm, wm, a_mid, hi, a_hi, H, W, f)) return rc;
    HP_REQUIRE(out, HP_ERR_NULL, "hp_fuse_multiscale: out is null");
    HP_REQUIRE(n_maps >= 0, HP_ERR_SHAPE, "hp_fuse_multiscale: n_maps=%d", n_maps);
    if (n_maps == 0) return HP_OK;
    const bool vec = (W % 4 == 0) && aligned16(out) && (!hi || aligned16(hi));
    const size_t total = static_cast<size_t>(n_maps) * H * W / (vec ? 4 : 1);
    size_t blocks = (total + 255) / 256;
    const size_t cap = 148u * 8u * 16u;
    const int grid = static_cast<int>(blocks < cap ? blocks : cap);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (vec) fuse_kernel<true><<<grid, 256, 0, s>>>(f, n_maps, out);
    else fuse_kernel<false><<<grid, 256, 0, s>>>(f, n_maps, out);
    return launch_status("hp_fuse_multiscale");
}

extern "C" HP_API int hp_fuse_decode_pck(const float* lo, int hl, int wl, float a_lo, const float* mid, int hm, int wm,
                                         float a_mid, const float* hi, float a_hi, const float* tgt_xy, int B, int K,
                                         int H, int W, double thr, float* pred_xy, float* maxvals, int32_t* counts,
                                         double* acc_out, void* workspace, hp_stream_t stream) {
    FuseSrc f;
    if (int rc = make_src("hp_fuse_decode_pck", lo, hl, wl, a_lo, mid, hm, wm, a_mid, hi, a_hi, H, W, f)) return rc;
    HP_REQUIRE(tgt_xy && pred_xy && acc_out && workspace, HP_ERR_NULL, "hp_fuse_decode_pck: null pointer");
    HP_REQUIRE(B > 0 && K > 0 && K <= HP_MAX_K, HP_ERR_SHAPE, "hp_fuse_decode_pck: bad B=%d K=%d", B, K);
    FuseDecodeLaunch l{f, tgt_xy, B * K, K, thr, pred_xy, maxvals, counts, acc_out, static_cast<Workspace*>(workspace),
                       static_cast<cudaStream_t>(stream)};
    dispatch_map_walk(H * W, !hi || aligned16(hi), l);
    return launch_status("hp_fuse_decode_pck");
}
