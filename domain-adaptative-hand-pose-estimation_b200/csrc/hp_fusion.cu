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
#include <cstdlib>

#include "hp_common.cuh"
#include "hp_decode.cuh"
#include "hp_dispatch.cuh"
#include "hp_tma.cuh"
#include "hp_peer_step.cuh"
#include "hp_bilinear_block.cuh"

namespace hp {


// warp-level argmax with the numpy tie rule (lowest index, NaN wins); every lane gets the result
__device__ __forceinline__ ArgMax warp_argmax_rows(ArgMax am, int lane) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        ArgMax b;
        b.v = __shfl_xor_sync(0xffffffffu, am.v, o);
        b.i = __shfl_xor_sync(0xffffffffu, am.i, o);
        am = ((lane & o) == 0) ? am_merge(am, b) : am_merge(b, am);
    }
    return am;
}

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


// ---------------------------------------------------------------------------------------------
// Row-walking shape (the production path for W in {4, 8, ..., 128}, W % 4 == 0):
// the first version above re-derived four bilinear taps and gathered eight source values for EVERY output
// element (~250 instructions per float4: 0.07-0.14 of HBM).  Bilinear interpolation is separable,
//     out(y, x) = l0y * T(i0y, x) + l1y * T(i1y, x),   T(r, x) = l0x * src[r][i0x] + l1x * src[r][i1x],
// so here a lane owns four output columns for a strip of consecutive rows: its column taps are computed once,
// the row taps once per block (shared-memory table), and the horizontally interpolated source rows T live in
// registers and are refreshed only when the source row changes (every 2nd / 4th output row at x2 / x4).
// Per output float4 that leaves: one 128-bit load of `hi` (the HBM stream), ~30 FMA-pipe instructions and the
// argmax bookkeeping.  The blend uses fused multiply-adds; the materialising kernel and the decode kernel
// share this code, so decoding in registers equals decoding the materialised map bit for bit.
// ---------------------------------------------------------------------------------------------
struct RowTap {
    int i0, i1;
    float l0, l1;
};
constexpr int kRowWarps = 4;

struct RowSource {  // one low-resolution source as seen by a lane (values in pairs: packed FMUL2 / FFMA2)
    const float* base;   // current map in global memory (rows kernel)
    uint32_t base_s;     // current map in shared memory (staged kernel), shared-window byte address
    int w;               // source width (elements)
    int off0[4], off1[4];        // element offsets of the two column taps of each output column
    float2 l0[2], l1[2];         // column weights, columns (0,1) and (2,3)
    int cur0, cur1;              // source rows held in t0 / t1 (-1: none)
    float2 t0[2], t1[2];
};

__device__ __forceinline__ void row_source_init(RowSource& s, int x0, float sx, int in_w, int out_w) {
    s.w = in_w;
    float l0[4], l1[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const Tap t = make_tap(x0 + c, sx, in_w, out_w);
        s.off0[c] = t.i0;
        s.off1[c] = t.i1;
        l0[c] = t.l0;
        l1[c] = t.l1;
    }
    s.l0[0] = make_float2(l0[0], l0[1]); s.l0[1] = make_float2(l0[2], l0[3]);
    s.l1[0] = make_float2(l1[0], l1[1]); s.l1[1] = make_float2(l1[2], l1[3]);
    s.cur0 = s.cur1 = -1;
    s.base = nullptr;
    s.base_s = 0;
}
template <bool STAGED>
__device__ __forceinline__ void row_source_load(const RowSource& s, int r, float2 (&t)[2]) {
    float a[4], b[4];
    if (STAGED) {  // the map sits in shared memory: 32-bit addresses, one add per load
        const uint32_t row = s.base_s + 4u * static_cast<uint32_t>(r * s.w);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            a[c] = lds_f32(row + 4u * static_cast<uint32_t>(s.off0[c]));
            b[c] = lds_f32(row + 4u * static_cast<uint32_t>(s.off1[c]));
        }
    } else {
        const float* row = s.base + r * s.w;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            a[c] = __ldg(row + s.off0[c]);
            b[c] = __ldg(row + s.off1[c]);
        }
    }
    t[0] = __ffma2_rn(s.l1[0], make_float2(b[0], b[1]), __fmul2_rn(s.l0[0], make_float2(a[0], a[1])));
    t[1] = __ffma2_rn(s.l1[1], make_float2(b[2], b[3]), __fmul2_rn(s.l0[1], make_float2(a[2], a[3])));
}
// vertical blend of the source at an output row with taps ty (refreshing the cached rows as needed)
template <bool STAGED>
__device__ __forceinline__ void row_source_at(RowSource& s, const RowTap ty, float2 (&v)[2]) {
    if (ty.i1 != s.cur1 || ty.i0 != s.cur0) {
        if (ty.i0 != s.cur0) {
            if (ty.i0 == s.cur1) {
                s.t0[0] = s.t1[0];
                s.t0[1] = s.t1[1];
            } else {
                row_source_load<STAGED>(s, ty.i0, s.t0);
            }
            s.cur0 = ty.i0;
        }
        if (ty.i1 != s.cur1) {
            if (ty.i1 == s.cur0) {
                s.t1[0] = s.t0[0];
                s.t1[1] = s.t0[1];
            } else {
                row_source_load<STAGED>(s, ty.i1, s.t1);
            }
            s.cur1 = ty.i1;
        }
    }
    const float2 l0 = make_float2(ty.l0, ty.l0), l1 = make_float2(ty.l1, ty.l1);
    v[0] = __ffma2_rn(l1, s.t1[0], __fmul2_rn(l0, s.t0[0]));
    v[1] = __ffma2_rn(l1, s.t1[1], __fmul2_rn(l0, s.t0[1]));
}

struct RowWalk {  // geometry of a block's walk over one map
    int cpr;      // float4 per output row
    int rps;      // output rows covered by a warp per step
    int strip;    // rows per warp
};

// fused float4 of output row `row` for this lane (hi4: the lane's four `hi` values or zeros)
template <bool STAGED = false>
__device__ __forceinline__ float4 fused_row4(const FuseSrc& f, RowSource& lo, RowSource& mid, const RowTap* s_rows, int row,
                                             float4 hi4) {
    float2 vl[2], r[2];
    row_source_at<STAGED>(lo, s_rows[row], vl);
    const float2 al = make_float2(f.a_lo, f.a_lo);
    r[0] = __fmul2_rn(al, vl[0]);
    r[1] = __fmul2_rn(al, vl[1]);
    if (f.mid) {
        float2 vm[2];
        row_source_at<STAGED>(mid, s_rows[f.H + row], vm);
        const float2 am = make_float2(f.a_mid, f.a_mid);
        r[0] = __ffma2_rn(am, vm[0], r[0]);
        r[1] = __ffma2_rn(am, vm[1], r[1]);
    }
    if (f.hi) {
        const float2 ah = make_float2(f.a_hi, f.a_hi);
        r[0] = __ffma2_rn(ah, make_float2(hi4.x, hi4.y), r[0]);
        r[1] = __ffma2_rn(ah, make_float2(hi4.z, hi4.w), r[1]);
    }
    return make_float4(r[0].x, r[0].y, r[1].x, r[1].y);
}

template <bool DECODE>
__global__ void __launch_bounds__(32 * kRowWarps)
    fuse_rows_kernel(const FuseSrc f, const RowWalk g, int n_maps, float* __restrict__ out, const float* __restrict__ tgt_xy,
                     int K, double thr, float* __restrict__ pred_xy, float* __restrict__ maxvals,
                     int32_t* __restrict__ counts_out, double* __restrict__ acc_out, Workspace* __restrict__ ws) {
    extern __shared__ RowTap s_rows[];  // [H] taps into lo, then [H] taps into mid
    __shared__ ArgMax s_am[kRowWarps];
    __shared__ int s_nan;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int HW = f.H * f.W;
    for (int r = threadIdx.x; r < f.H; r += blockDim.x) {
        const Tap tl = make_tap(r, f.sy_lo, f.hl, f.H);
        s_rows[r] = RowTap{tl.i0, tl.i1, tl.l0, tl.l1};
        if (f.mid) {
            const Tap tm = make_tap(r, f.sy_mid, f.hm, f.H);
            s_rows[f.H + r] = RowTap{tm.i0, tm.i1, tm.l0, tm.l1};
        }
    }
    const int x0 = (lane % g.cpr) * 4, ro = lane / g.cpr;
    RowSource lo, mid;
    row_source_init(lo, x0, f.sx_lo, f.wl, f.W);
    row_source_init(mid, x0, f.mid ? f.sx_mid : 1.0f, f.mid ? f.wm : f.W, f.W);
    const int row_begin = warp * g.strip + ro, row_end = min(f.H, (warp + 1) * g.strip);
    __syncthreads();

    for (int map = blockIdx.x; map < n_maps; map += gridDim.x) {
        lo.base = f.lo + static_cast<size_t>(map) * f.hl * f.wl;
        lo.cur0 = lo.cur1 = -1;
        if (f.mid) {
            mid.base = f.mid + static_cast<size_t>(map) * f.hm * f.wm;
            mid.cur0 = mid.cur1 = -1;
        }
        const float* hi = f.hi ? f.hi + static_cast<size_t>(map) * HW + x0 : nullptr;
        float* o = DECODE ? nullptr : out + static_cast<size_t>(map) * HW + x0;
        float best = -INFINITY, witness = 0.0f;
        int best_row = row_begin;
        for (int row = row_begin; row < row_end; row += 4 * g.rps) {
            float4 h[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {  // four rows of the HBM stream in flight
                const int r = row + u * g.rps;
                h[u] = (hi && r < row_end) ? ldg_stream4(reinterpret_cast<const float4*>(hi + r * f.W)) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int r = row + u * g.rps;
                if (r < row_end) {
                    const float4 v = fused_row4(f, lo, mid, s_rows, r, h[u]);
                    if (DECODE) {
                        const float m4 = fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w));
                        best_row = (m4 > best) ? r : best_row;  // strict: the earlier row keeps ties
                        best = fmaxf(best, m4);
                        witness += (v.x + v.y) + (v.z + v.w);    // NaN / inf-inf witness
                    } else {
                        stg_stream4(reinterpret_cast<float4*>(o + r * f.W), v);
                    }
                }
            }
        }
        if (DECODE) {
            // the lane's first maximum: recompute its row (same code, same bits) and pick the component
            ArgMax am = am_init();
            if (row_begin < row_end) {
                const float4 hb = hi ? ldg_stream4(reinterpret_cast<const float4*>(hi + best_row * f.W)) : make_float4(0.f, 0.f, 0.f, 0.f);
                const float4 v = fused_row4(f, lo, mid, s_rows, best_row, hb);
                const int comp = (v.x == best) ? 0 : ((v.y == best) ? 1 : ((v.z == best) ? 2 : 3));
                am.v = best;
                am.i = best_row * f.W + x0 + comp;
            }
            am = warp_argmax_rows(am, lane);
            const bool bad = __any_sync(0xffffffffu, witness != witness);
            if (threadIdx.x == 0) s_nan = 0;
            __syncthreads();
            if (lane == 0) {
                s_am[warp] = am;
                if (bad) s_nan = 1;
            }
            __syncthreads();
            if (s_nan) {
                // a NaN (or +inf with -inf) somewhere in the fused map: element-wise scan with numpy's exact rules
                ArgMax sx = am_init();
                for (int r = row_begin; r < row_end; r += g.rps) {
                    const float4 hb = hi ? ldg_stream4(reinterpret_cast<const float4*>(hi + r * f.W)) : make_float4(0.f, 0.f, 0.f, 0.f);
                    am_scan4<true>(sx, fused_row4(f, lo, mid, s_rows, r, hb), r * f.W + x0);
                }
                sx = warp_argmax_rows(sx, lane);
                __syncthreads();
                if (lane == 0) s_am[warp] = sx;
                __syncthreads();
            }
            if (threadIdx.x == 0) {
                ArgMax a = s_am[0];
#pragma unroll
                for (int w = 1; w < kRowWarps; ++w) a = am_merge(a, s_am[w]);
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
            __syncthreads();  // s_am / s_nan are reused by the next map
        }
    }
    if (DECODE) {
        if (last_block_arrives(&ws->counter, gridDim.x)) pck_publish(ws, K, counts_out, acc_out);
    }
}

// rows kernel applicable?  W a multiple of 4 with W/4 dividing 32, 16-byte aligned streams, table fits
static bool rows_geometry(const FuseSrc& f, const float* out, RowWalk& g) {
    if (f.W % 4 != 0 || f.W > 128 || (32 % (f.W / 4)) != 0) return false;
    if (f.hi && !aligned16(f.hi)) return false;
    if (out && !aligned16(out)) return false;
    if (f.H > 1024) return false;
    g.cpr = f.W / 4;
    g.rps = 32 / g.cpr;
    const int steps = (f.H + kRowWarps * g.rps - 1) / (kRowWarps * g.rps);
    g.strip = steps * g.rps;
    return true;
}
static int rows_grid(int n_maps) {
    int sms = hp_device_sm_count();
    if (sms <= 0) sms = 148;
    const int cap = sms * 4;  // 4 blocks of 4 warps per SM (128 registers per thread): persistent, grid-stride over the maps
    return n_maps < cap ? n_maps : cap;
}


// ---------------------------------------------------------------------------------------------
// Staged shape: the same row walk, but the low-resolution maps are copied into shared memory by the copy
// engine (cp.async.bulk + mbarrier), double-buffered per block, so refreshing an interpolated row is a
// handful of shared-memory loads instead of L2 round trips, and the next map's sources arrive while the
// current one is processed.  No block barrier in the loop: a warp that finishes its strip of a map takes a
// ticket from a shared counter; the LAST warp of a map merges the four partial argmaxes, scores PCK and
// re-fills the buffer with the sources of the block's map after next (everybody has read it by then).
// ---------------------------------------------------------------------------------------------
struct StagedCtl {
    unsigned long long full[2];  // mbarriers: sources of buffer b have landed
    int done[2];                 // warps that have finished the map in buffer b
    int nan[2];
    ArgMax am[2][kRowWarps];
};

template <bool DECODE>
__global__ void __launch_bounds__(32 * kRowWarps, 4)
    fuse_staged_kernel(const FuseSrc f, const RowWalk g, int n_maps, float* __restrict__ out, const float* __restrict__ tgt_xy,
                       int K, double thr, float* __restrict__ pred_xy, float* __restrict__ maxvals,
                       int32_t* __restrict__ counts_out, double* __restrict__ acc_out, Workspace* __restrict__ ws) {
    extern __shared__ __align__(128) unsigned char s_raw[];
    __shared__ StagedCtl ctl;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int HW = f.H * f.W;
    const int lo_elems = f.hl * f.wl, mid_elems = f.mid ? f.hm * f.wm : 0;
    const uint32_t lo_bytes = 4u * lo_elems, mid_bytes = 4u * mid_elems, buf_bytes = lo_bytes + mid_bytes;
    float* s_src = reinterpret_cast<float*>(s_raw);                               // [2][lo | mid]
    RowTap* s_rows = reinterpret_cast<RowTap*>(s_raw + 2 * static_cast<size_t>(buf_bytes));  // [H] lo taps, [H] mid taps
    const int n_local = (n_maps - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
    const uint32_t full_u32 = smem_addr(&ctl.full[0]), src_u32 = smem_addr(s_src);
    const uint64_t pol = l2_evict_first_policy();

    auto request = [&](int j) {  // one thread: sources of the block's j-th map -> buffer j & 1
        const int b = j & 1;
        const size_t map = static_cast<size_t>(blockIdx.x) + static_cast<size_t>(j) * gridDim.x;
        mbar_arrive_expect_tx(full_u32 + 8 * b, buf_bytes);
        bulk_load(src_u32 + b * buf_bytes, f.lo + map * lo_elems, lo_bytes, full_u32 + 8 * b, pol);
        if (f.mid) bulk_load(src_u32 + b * buf_bytes + lo_bytes, f.mid + map * mid_elems, mid_bytes, full_u32 + 8 * b, pol);
    };
    if (threadIdx.x == 0) {
        mbar_init(full_u32, 1);
        mbar_init(full_u32 + 8, 1);
        mbar_init_fence();
        ctl.done[0] = ctl.done[1] = 0;
        ctl.nan[0] = ctl.nan[1] = 0;
        if (n_local > 0) request(0);
        if (n_local > 1) request(1);
    }
    for (int r = threadIdx.x; r < f.H; r += blockDim.x) {
        const Tap tl = make_tap(r, f.sy_lo, f.hl, f.H);
        s_rows[r] = RowTap{tl.i0, tl.i1, tl.l0, tl.l1};
        if (f.mid) {
            const Tap tm = make_tap(r, f.sy_mid, f.hm, f.H);
            s_rows[f.H + r] = RowTap{tm.i0, tm.i1, tm.l0, tm.l1};
        }
    }
    const int x0 = (lane % g.cpr) * 4, ro = lane / g.cpr;
    RowSource lo, mid;
    row_source_init(lo, x0, f.sx_lo, f.wl, f.W);
    row_source_init(mid, x0, f.mid ? f.sx_mid : 1.0f, f.mid ? f.wm : f.W, f.W);
    const int row_begin = warp * g.strip + ro, row_end = min(f.H, (warp + 1) * g.strip);
    __syncthreads();  // the only block barrier: control block and tap table are set up

    for (int j = 0; j < n_local; ++j) {
        const int b = j & 1;
        const int map = static_cast<int>(blockIdx.x) + j * static_cast<int>(gridDim.x);
        lo.base_s = src_u32 + b * buf_bytes;
        mid.base_s = lo.base_s + lo_bytes;
        lo.cur0 = lo.cur1 = mid.cur0 = mid.cur1 = -1;
        const float* hi = f.hi ? f.hi + static_cast<size_t>(map) * HW + x0 : nullptr;
        float* o = DECODE ? nullptr : out + static_cast<size_t>(map) * HW + x0;
        float best = -INFINITY, witness = 0.0f;
        int best_row = row_begin;
        float4 h[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {  // the first rows of the HBM stream are requested before the sources are awaited
            const int r = row_begin + u * g.rps;
            h[u] = (hi && r < row_end) ? ldg_stream4(reinterpret_cast<const float4*>(hi + r * f.W)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        mbar_wait(full_u32 + 8 * b, static_cast<uint32_t>(j >> 1) & 1u);
        for (int row = row_begin; row < row_end; row += 4 * g.rps) {
            float4 hn[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {  // next step's rows in flight while this step is blended
                const int r = row + (4 + u) * g.rps;
                hn[u] = (hi && r < row_end) ? ldg_stream4(reinterpret_cast<const float4*>(hi + r * f.W)) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int r = row + u * g.rps;
                if (r < row_end) {
                    const float4 v = fused_row4<true>(f, lo, mid, s_rows, r, h[u]);
                    if (DECODE) {
                        const float m4 = fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w));
                        best_row = (m4 > best) ? r : best_row;  // strict: the earlier row keeps ties
                        best = fmaxf(best, m4);
                        witness += (v.x + v.y) + (v.z + v.w);    // NaN / inf-inf witness
                    } else {
                        stg_stream4(reinterpret_cast<float4*>(o + r * f.W), v);
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) h[u] = hn[u];
        }
        ArgMax am = am_init();
        bool bad = false;
        if (DECODE) {
            if (row_begin < row_end) {  // the lane's first maximum: recompute its row (same code, same bits)
                const float4 hb = hi ? ldg_stream4(reinterpret_cast<const float4*>(hi + best_row * f.W)) : make_float4(0.f, 0.f, 0.f, 0.f);
                const float4 v = fused_row4<true>(f, lo, mid, s_rows, best_row, hb);
                const int comp = (v.x == best) ? 0 : ((v.y == best) ? 1 : ((v.z == best) ? 2 : 3));
                am.v = best;
                am.i = best_row * f.W + x0 + comp;
            }
            am = warp_argmax_rows(am, lane);
            bad = __any_sync(0xffffffffu, witness != witness);
        }
        // ---- ticket: the last warp of this map closes it and re-fills the buffer ------------------------------
        __syncwarp();
        int last = 0;
        if (lane == 0) {
            if (DECODE) {
                ctl.am[b][warp] = am;
                if (bad) atomicOr(&ctl.nan[b], 1);
            }
            __threadfence_block();
            last = (atomicAdd(&ctl.done[b], 1) == kRowWarps - 1) ? 1 : 0;
            if (last) __threadfence_block();
        }
        last = __shfl_sync(0xffffffffu, last, 0);
        if (last) {
            if (DECODE) {
                ArgMax a = ctl.am[b][0];
#pragma unroll
                for (int w = 1; w < kRowWarps; ++w) a = am_merge(a, ctl.am[b][w]);
                if (*reinterpret_cast<volatile int*>(&ctl.nan[b])) {
                    // a NaN (or +inf with -inf) somewhere in the fused map: this warp rescans the whole map
                    // element-wise with numpy's exact rules (the sources are still staged)
                    ArgMax sx = am_init();
                    lo.cur0 = lo.cur1 = mid.cur0 = mid.cur1 = -1;
                    for (int r = ro; r < f.H; r += g.rps) {
                        const float4 hb = hi ? ldg_stream4(reinterpret_cast<const float4*>(hi + r * f.W)) : make_float4(0.f, 0.f, 0.f, 0.f);
                        am_scan4<true>(sx, fused_row4<true>(f, lo, mid, s_rows, r, hb), r * f.W + x0);
                    }
                    a = warp_argmax_rows(sx, lane);
                }
                if (lane == 0) {
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
            __syncwarp();
            if (lane == 0) {
                ctl.done[b] = 0;
                ctl.nan[b] = 0;
                __threadfence_block();
                if (j + 2 < n_local) request(j + 2);  // every warp has finished reading buffer b
            }
        }
    }
    if (DECODE) {
        if (last_block_arrives(&ws->counter, gridDim.x)) pck_publish(ws, K, counts_out, acc_out);
    }
}

// staged kernel applicable?  rows geometry + bulk-copy alignment + both buffers and the tap table fit
static bool staged_geometry(const FuseSrc& f, const float* out, RowWalk& g, size_t& smem) {
    if (!rows_geometry(f, out, g)) return false;
    const size_t lo_b = 4ull * f.hl * f.wl, mid_b = f.mid ? 4ull * f.hm * f.wm : 0;
    if (lo_b % 16 != 0 || mid_b % 16 != 0 || !aligned16(f.lo) || (f.mid && !aligned16(f.mid))) return false;
    smem = 2 * (lo_b + mid_b) + sizeof(RowTap) * 2 * static_cast<size_t>(f.H);
    return smem <= 48 * 1024;  // 4 blocks per SM
}

}  // namespace hp
#include "hp_fusion_block.cuh"
namespace hp {

// HP_FUSE_SHAPE=r forces the row-walking kernels (A/B comparisons; the block kernel is the default for exact scales)
static bool fuse_rows_forced() {
    const char* e = getenv("HP_FUSE_SHAPE");
    return e != nullptr && e[0] == 'r';
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
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    {
        BlockWalk bg;
        int sl = 0, sm = 0, n_warps = 0;
        size_t smem = 0;
        if (!fuse_rows_forced() && block_geometry(f, out, bg, sl, sm, n_warps, smem)) {
            launch_fuse_block<false>(f, bg, sl, sm, n_warps, smem, n_maps, out, nullptr, 0, 0.0, nullptr, nullptr, nullptr, nullptr,
                                     nullptr, s);
            return launch_status("hp_fuse_multiscale");
        }
    }
    RowWalk g;
    size_t staged_smem = 0;
    if (staged_geometry(f, out, g, staged_smem)) {
        fuse_staged_kernel<false><<<rows_grid(n_maps), 32 * kRowWarps, staged_smem, s>>>(f, g, n_maps, out, nullptr, 0, 0.0, nullptr,
                                                                                       nullptr, nullptr, nullptr, nullptr);
        return launch_status("hp_fuse_multiscale");
    }
    if (rows_geometry(f, out, g)) {
        const size_t smem = sizeof(RowTap) * 2 * static_cast<size_t>(H);
        fuse_rows_kernel<false><<<rows_grid(n_maps), 32 * kRowWarps, smem, s>>>(f, g, n_maps, out, nullptr, 0, 0.0, nullptr,
                                                                                nullptr, nullptr, nullptr, nullptr);
        return launch_status("hp_fuse_multiscale");
    }
    const bool vec = (W % 4 == 0) && aligned16(out) && (!hi || aligned16(hi));
    const size_t total = static_cast<size_t>(n_maps) * H * W / (vec ? 4 : 1);
    size_t blocks = (total + 255) / 256;
    const size_t cap = 148u * 8u * 16u;
    const int grid = static_cast<int>(blocks < cap ? blocks : cap);
    if (vec) fuse_kernel<true><<<grid, 256, 0, s>>>(f, n_maps, out);
    else fuse_kernel<false><<<grid, 256, 0, s>>>(f, n_maps, out);
    return launch_status("hp_fuse_multiscale");
}

/* train1.py:410-424 in one call: out [n_maps, H, W] = a_lo * up(lo) + a_mid * up(mid) (`target5`) and
 * out2 [n_maps, H2, W2] = a_lo2 * up(lo) (`target0`), both nn.Upsample(mode='bilinear').  ONE launch for the driver's geometry
 * (lo x4 and mid x2 -> H x W, lo x2 -> H2 x W2 = H/2 x W/2); anything else: two hp_fuse_multiscale launches. */
extern "C" HP_API int hp_fuse_multiscale_pair(const float* lo, int hl, int wl, float a_lo, const float* mid, int hm, int wm, float a_mid,
                                              int n_maps, int H, int W, float* out, float a_lo2, int H2, int W2, float* out2,
                                              hp_stream_t stream) {
    HP_REQUIRE(out && out2, HP_ERR_NULL, "hp_fuse_multiscale_pair: null output");
    if (n_maps > 0 && mid && H2 * 2 == H && W2 * 2 == W && !fuse_rows_forced()) {
        FuseSrc f;
        if (int rc = make_src("hp_fuse_multiscale_pair", lo, hl, wl, a_lo, mid, hm, wm, a_mid, nullptr, 0.0f, H, W, f)) return rc;
        BlockWalk bg;
        int sl = 0, sm = 0, n_warps = 0;
        size_t smem = 0;
        if (block_geometry(f, out, bg, sl, sm, n_warps, smem) &&
            launch_fuse_pair(f, bg, sl, sm, n_warps, smem, n_maps, out, out2, a_lo2, static_cast<cudaStream_t>(stream)))
            return launch_status("hp_fuse_multiscale_pair");
    }
    if (int rc = hp_fuse_multiscale(lo, hl, wl, a_lo, mid, hm, wm, a_mid, nullptr, 0.0f, n_maps, H, W, out, stream)) return rc;
    return hp_fuse_multiscale(lo, hl, wl, a_lo2, nullptr, 0, 0, 0.0f, nullptr, 0.0f, n_maps, H2, W2, out2, stream);
}

// link != nullptr: a sharded call; *exchanged tells whether the kernel summed the counts over the ranks itself
static int fuse_decode_pck_impl(const float* lo, int hl, int wl, float a_lo, const float* mid, int hm, int wm,
                                float a_mid, const float* hi, float a_hi, const float* tgt_xy, int B, int K,
                                int H, int W, double thr, float* pred_xy, float* maxvals, int32_t* counts,
                                double* acc_out, void* workspace, hp_stream_t stream, const PeerLink* link, bool* exchanged,
                                int defer = 0, long long* partial_out = nullptr, double* result_out = nullptr);

extern "C" HP_API int hp_fuse_decode_pck(const float* lo, int hl, int wl, float a_lo, const float* mid, int hm, int wm,
                                         float a_mid, const float* hi, float a_hi, const float* tgt_xy, int B, int K,
                                         int H, int W, double thr, float* pred_xy, float* maxvals, int32_t* counts,
                                         double* acc_out, void* workspace, hp_stream_t stream) {
    return fuse_decode_pck_impl(lo, hl, wl, a_lo, mid, hm, wm, a_mid, hi, a_hi, tgt_xy, B, K, H, W, thr, pred_xy, maxvals, counts,
                                acc_out, workspace, stream, nullptr, nullptr);
}

/* The sharded form (configs[3] on the GPUs of one node): the same step on this rank's samples; `counts` / `acc_out` then hold
 * the totals over ALL ranks (keypoint_detection.py:63-92 on the concatenated batch).  Where the staged kernel applies
 * (32 / 64 / 128) its last block exchanges the 2K integer counts over the peer mailboxes itself; otherwise a one-warp
 * hp_pck_finalize_peer launch follows.  `mailboxes`: as for hp_pipeline_fused_peer. */
extern "C" HP_API int hp_fuse_decode_pck_peer(const float* lo, int hl, int wl, float a_lo, const float* mid, int hm, int wm,
                                              float a_mid, const float* hi, float a_hi, const float* tgt_xy, int B, int K,
                                              int H, int W, double thr, float* pred_xy, float* maxvals, int32_t* counts,
                                              double* acc_out, void* workspace, void* const* mailboxes, int rank, int world,
                                              unsigned int flags, int64_t* partial_out, double* result_out, hp_stream_t stream) {
    HP_REQUIRE(counts && mailboxes, HP_ERR_NULL, "hp_fuse_decode_pck_peer: null pointer");
    const int defer = (flags & 2u) != 0 && world > 1;  // HP_PIPE_DEFER_EXCHANGE
    HP_REQUIRE(!defer || (partial_out && result_out), HP_ERR_NULL, "hp_fuse_decode_pck_peer: a deferred step needs partial_out and result_out");
    HP_REQUIRE(world > 0 && world <= kPeerMaxWorld && rank >= 0 && rank < world && K > 0 && K <= HP_MAX_K && peer_shape_ok(K, world),
               HP_ERR_ARG, "hp_fuse_decode_pck_peer: rank=%d world=%d K=%d (K <= 27 when sharded)", rank, world, K);
    PeerLink link{};
    for (int r = 0; r < world; ++r) {
        HP_REQUIRE(mailboxes[r], HP_ERR_NULL, "hp_fuse_decode_pck_peer: mailbox %d is null", r);
        link.mailbox[r] = static_cast<unsigned long long*>(mailboxes[r]);
    }
    link.rank = rank;
    link.world = world;
    bool exchanged = false;
    if (int rc = fuse_decode_pck_impl(lo, hl, wl, a_lo, mid, hm, wm, a_mid, hi, a_hi, tgt_xy, B, K, H, W, thr, pred_xy, maxvals,
                                      counts, acc_out, workspace, stream, world > 1 ? &link : nullptr, &exchanged, defer,
                                      reinterpret_cast<long long*>(partial_out), result_out))
        return rc;
    if (world > 1 && !exchanged) {
        HP_REQUIRE(!defer, HP_ERR_SHAPE, "hp_fuse_decode_pck_peer: only the staged 32/64/128 kernel can defer its exchange");
        return hp_pck_finalize_peer(counts, mailboxes, rank, world, K, counts, acc_out, stream);
    }
    return HP_OK;
}

static int fuse_decode_pck_impl(const float* lo, int hl, int wl, float a_lo, const float* mid, int hm, int wm,
                                float a_mid, const float* hi, float a_hi, const float* tgt_xy, int B, int K,
                                int H, int W, double thr, float* pred_xy, float* maxvals, int32_t* counts,
                                double* acc_out, void* workspace, hp_stream_t stream, const PeerLink* link, bool* exchanged,
                                int defer, long long* partial_out, double* result_out) {
    FuseSrc f;
    if (int rc = make_src("hp_fuse_decode_pck", lo, hl, wl, a_lo, mid, hm, wm, a_mid, hi, a_hi, H, W, f)) return rc;
    HP_REQUIRE(tgt_xy && pred_xy && acc_out && workspace, HP_ERR_NULL, "hp_fuse_decode_pck: null pointer");
    HP_REQUIRE(B > 0 && K > 0 && K <= HP_MAX_K, HP_ERR_SHAPE, "hp_fuse_decode_pck: bad B=%d K=%d", B, K);
    {
        BlockWalk bg;
        int sl = 0, sm = 0, n_warps = 0;
        size_t smem = 0;
        if (!fuse_rows_forced() && block_geometry(f, nullptr, bg, sl, sm, n_warps, smem)) {
            const bool ex = launch_fuse_block<true>(f, bg, sl, sm, n_warps, smem, B * K, nullptr, tgt_xy, K, thr, pred_xy, maxvals, counts,
                                                    acc_out, static_cast<Workspace*>(workspace), static_cast<cudaStream_t>(stream), link,
                                                    defer, partial_out, result_out);
            if (exchanged) *exchanged = ex;
            return launch_status("hp_fuse_decode_pck");
        }
    }
    RowWalk g;
    size_t staged_smem = 0;
    if (staged_geometry(f, nullptr, g, staged_smem)) {
        fuse_staged_kernel<true><<<rows_grid(B * K), 32 * kRowWarps, staged_smem, static_cast<cudaStream_t>(stream)>>>(
            f, g, B * K, nullptr, tgt_xy, K, thr, pred_xy, maxvals, counts, acc_out, static_cast<Workspace*>(workspace));
        return launch_status("hp_fuse_decode_pck");
    }
    if (rows_geometry(f, nullptr, g)) {
        const size_t smem = sizeof(RowTap) * 2 * static_cast<size_t>(H);
        fuse_rows_kernel<true><<<rows_grid(B * K), 32 * kRowWarps, smem, static_cast<cudaStream_t>(stream)>>>(
            f, g, B * K, nullptr, tgt_xy, K, thr, pred_xy, maxvals, counts, acc_out, static_cast<Workspace*>(workspace));
        return launch_status("hp_fuse_decode_pck");
    }
    FuseDecodeLaunch l{f, tgt_xy, B * K, K, thr, pred_xy, maxvals, counts, acc_out, static_cast<Workspace*>(workspace),
                       static_cast<cudaStream_t>(stream)};
    dispatch_map_walk(H * W, !hi || aligned16(hi), l);
    return launch_status("hp_fuse_decode_pck");
}
