// hp_common.cuh - device helpers shared by every kernel of libhp_b200.so (sm_100a).
//
// Data layout: heatmaps are contiguous NCHW fp32; one "map" = H*W floats of one (sample, joint).
// All hot kernels are HBM-bound streaming reductions, so the helpers are about
//   * 128-bit read-once loads that bypass L1 allocation,
//   * keeping a whole map (or a tile of it) in registers so it is read from HBM exactly once,
//   * warp-shuffle + shared-memory reductions of small stat structs,
//   * numpy/torch-exact semantics for argmax (first index, NaN wins) and float64 PCK.
#pragma once

#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/hp_b200.h"

namespace hp {

// ---------------------------------------------------------------------------------------------
// host-side error plumbing (defined in hp_api.cu)
// ---------------------------------------------------------------------------------------------
int fail(int code, const char* fmt, ...);
int launch_status(const char* what);

#define HP_REQUIRE(cond, code, ...)                   \
    do {                                              \
        if (!(cond)) return ::hp::fail((code), __VA_ARGS__); \
    } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline bool aligned8(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 7u) == 0; }
inline bool aligned4(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 3u) == 0; }

// ---------------------------------------------------------------------------------------------
// exact unsigned division by a runtime constant (row/col from a flat index without IDIV)
// ---------------------------------------------------------------------------------------------
struct FastDiv {
    uint32_t d, m, s1, s2;
    FastDiv() : d(1), m(1), s1(0), s2(0) {}
    explicit FastDiv(uint32_t div) : d(div) {
        uint32_t l = 0;
        while ((1ull << l) < div) ++l;
        m = static_cast<uint32_t>(((1ull << 32) * ((1ull << l) - div)) / div + 1);
        s1 = l < 1 ? l : 1;
        s2 = l > 1 ? l - 1 : 0;
    }
    __device__ __forceinline__ uint32_t div(uint32_t n) const {
        uint32_t t = __umulhi(m, n);
        return (t + ((n - t) >> s1)) >> s2;
    }
    __device__ __forceinline__ void divmod(uint32_t n, uint32_t& q, uint32_t& r) const {
        q = div(n);
        r = n - q * d;
    }
};

// ---------------------------------------------------------------------------------------------
// loads
// ---------------------------------------------------------------------------------------------
// read-once 128-bit load: non-coherent path, no L1 allocation (the map is consumed from registers)
__device__ __forceinline__ float4 ldg_stream4(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ float ldg_stream1(const float* p) {
    float r;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ void stg_stream4(float4* p, float4 v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y),
                 "f"(v.z), "f"(v.w)
                 : "memory");
}

// How a map is walked.  EXACT: HW == TPM*NV*4 and 16-byte aligned -> one unguarded register tile.
// VEC: HW % 4 == 0 and aligned -> guarded float4 tiles.  SCALAR: anything (odd widths, offsets).
enum WalkMode { WALK_EXACT = 0, WALK_VEC = 1, WALK_SCALAR = 2 };

// Load tile `tile` of a map into registers.  Element e of the map sits in
//   v[j].{x,y,z,w}  with  e = tile*TPM*NV*4 + (j*TPM + t)*4 + c ;  out-of-range elements get `fill`.
template <int TPM, int NV, int MODE>
__device__ __forceinline__ void load_tile(const float* __restrict__ map, int HW, int tile, int t, float fill,
                                          float4 (&v)[NV]) {
    const int base = tile * (TPM * NV * 4);
    if (MODE == WALK_EXACT) {
        const float4* m4 = reinterpret_cast<const float4*>(map);
#pragma unroll
        for (int j = 0; j < NV; ++j) v[j] = ldg_stream4(m4 + j * TPM + t);
    } else if (MODE == WALK_VEC) {
        const float4* m4 = reinterpret_cast<const float4*>(map + base);
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            const int e = base + (j * TPM + t) * 4;
            v[j] = (e < HW) ? ldg_stream4(m4 + j * TPM + t) : make_float4(fill, fill, fill, fill);
        }
    } else {
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            const int e = base + (j * TPM + t) * 4;
            v[j].x = (e + 0 < HW) ? ldg_stream1(map + e + 0) : fill;
            v[j].y = (e + 1 < HW) ? ldg_stream1(map + e + 1) : fill;
            v[j].z = (e + 2 < HW) ? ldg_stream1(map + e + 2) : fill;
            v[j].w = (e + 3 < HW) ? ldg_stream1(map + e + 3) : fill;
        }
    }
}

template <int TPM, int NV>
__host__ __device__ __forceinline__ int tiles_for(int HW) {
    return (HW + TPM * NV * 4 - 1) / (TPM * NV * 4);
}

// ---------------------------------------------------------------------------------------------
// argmax with numpy semantics: first (lowest) index among equals, NaN beats everything
// ---------------------------------------------------------------------------------------------
struct ArgMax {
    float v;
    int i;
};
__device__ __forceinline__ ArgMax am_init() { return ArgMax{-INFINITY, 0x7fffffff}; }

__device__ __forceinline__ bool am_better(float v1, int i1, float v2, int i2) {
    const bool n1 = v1 != v1, n2 = v2 != v2;
    if (n1 | n2) return (n1 & n2) ? (i1 < i2) : n1;
    if (v1 != v2) return v1 > v2;
    return i1 < i2;
}
__device__ __forceinline__ ArgMax am_merge(ArgMax a, ArgMax b) { return am_better(b.v, b.i, a.v, a.i) ? b : a; }

// per-thread scan in increasing index order.  FAST ignores NaN (caller detects it and rescans).
template <bool NANSAFE>
__device__ __forceinline__ void am_scan1(ArgMax& a, float x, int idx) {
    bool take;
    if (NANSAFE) take = (x > a.v) || ((x != x) && (a.v == a.v));
    else take = x > a.v;
    if (take) {
        a.v = x;
        a.i = idx;
    }
}
template <bool NANSAFE>
__device__ __forceinline__ void am_scan4(ArgMax& a, float4 x, int idx0) {
    am_scan1<NANSAFE>(a, x.x, idx0);
    am_scan1<NANSAFE>(a, x.y, idx0 + 1);
    am_scan1<NANSAFE>(a, x.z, idx0 + 2);
    am_scan1<NANSAFE>(a, x.w, idx0 + 3);
}

// ---------------------------------------------------------------------------------------------
// online softmax pair (m, s): s = sum exp(x - m)
// ---------------------------------------------------------------------------------------------
constexpr float kLog2e = 1.4426950408889634f;

// warp-wide float maximum in one instruction (sm_100a; SASS CREDUX.MAX.F32); NaN inputs are ignored
__device__ __forceinline__ float warp_max_f32(float x) {
    float r;
    asm("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(x));
    return r;
}

// single MUFU instructions (SASS MUFU.EX2 / MUFU.LG2), ~2^-22 relative error
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
constexpr float kLn2 = 0.6931471805599453f;

__device__ __forceinline__ float exp_diff(float a, float b) {  // exp(a - b), 0 when a == -inf
    return (a == -INFINITY) ? 0.0f : exp2f((a - b) * kLog2e);
}
__device__ __forceinline__ void sm_merge(float& m, float& s, float m2, float s2) {
    const float mn = fmaxf(m, m2);
    s = s * exp_diff(m, mn) + s2 * exp_diff(m2, mn);
    m = mn;
}

// ---------------------------------------------------------------------------------------------
// Stat bundle reduced once per map: argmax pair, softmax pair, NSUM plain sums, one max, a flag
// ---------------------------------------------------------------------------------------------
template <int NSUM>
struct Stats {
    ArgMax am;
    float m, s;
    float sum[NSUM > 0 ? NSUM : 1];
    float mx;    // plain NaN-propagating max (ground-false normaliser)
    float flag;  // non-finite witness: fma(x, 0, flag) -> NaN once any NaN/inf was seen
};

template <int NSUM>
__device__ __forceinline__ void stats_init(Stats<NSUM>& a) {
    a.am = am_init();
    a.m = -INFINITY;
    a.s = 0.0f;
#pragma unroll
    for (int i = 0; i < NSUM; ++i) a.sum[i] = 0.0f;
    a.mx = -INFINITY;
    a.flag = 0.0f;
}

__device__ __forceinline__ float nanmax(float a, float b) {  // torch.max semantics: NaN propagates
    return (a != a) ? a : ((b != b) ? b : fmaxf(a, b));
}

template <int NSUM, bool AM, bool SM, bool MX>
__device__ __forceinline__ void stats_merge(Stats<NSUM>& a, const Stats<NSUM>& b) {
    if (AM) a.am = am_merge(a.am, b.am);
    if (SM) sm_merge(a.m, a.s, b.m, b.s);
#pragma unroll
    for (int i = 0; i < NSUM; ++i) a.sum[i] += b.sum[i];
    if (MX) a.mx = nanmax(a.mx, b.mx);
    a.flag += b.flag;
}

template <int NSUM, bool AM, bool SM, bool MX>
__device__ __forceinline__ Stats<NSUM> stats_shfl_xor(const Stats<NSUM>& a, int lane_mask) {
    Stats<NSUM> b;
    if (AM) {
        b.am.v = __shfl_xor_sync(0xffffffffu, a.am.v, lane_mask);
        b.am.i = __shfl_xor_sync(0xffffffffu, a.am.i, lane_mask);
    }
    if (SM) {
        b.m = __shfl_xor_sync(0xffffffffu, a.m, lane_mask);
        b.s = __shfl_xor_sync(0xffffffffu, a.s, lane_mask);
    }
#pragma unroll
    for (int i = 0; i < NSUM; ++i) b.sum[i] = __shfl_xor_sync(0xffffffffu, a.sum[i], lane_mask);
    if (MX) b.mx = __shfl_xor_sync(0xffffffffu, a.mx, lane_mask);
    b.flag = __shfl_xor_sync(0xffffffffu, a.flag, lane_mask);
    return b;
}

// Reduce over the TPM threads that share one map; every thread of the group gets the result.
//   TPM <= 32 : sub-warp butterfly (TPM must divide 32; groups are lane-aligned), no barrier.
//   TPM  > 32 : the group is the whole block (blockDim.x == TPM); `scratch` holds TPM/32 + 1 entries.
// The butterfly merges (a, b) and (b, a) on partner lanes; every merge op here is commutative up to
// the argmax/NaN tie rules, which are themselves symmetric, so all lanes agree bit-for-bit.
template <int TPM, int NSUM, bool AM, bool SM, bool MX>
__device__ __forceinline__ void group_reduce(Stats<NSUM>& a, Stats<NSUM>* scratch) {
    constexpr int W = TPM < 32 ? TPM : 32;
#pragma unroll
    for (int o = W / 2; o > 0; o >>= 1) {
        Stats<NSUM> b = stats_shfl_xor<NSUM, AM, SM, MX>(a, o);
        // keep the merge order identical on both partners: lower lane's value first
        const bool low = (threadIdx.x & o) == 0;
        Stats<NSUM> lo = low ? a : b, hi = low ? b : a;
        stats_merge<NSUM, AM, SM, MX>(lo, hi);
        a = lo;
    }
    if (TPM > 32) {
        constexpr int NW = TPM / 32;
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        __syncthreads();  // scratch may still be read from the previous reduction
        if (lane == 0) scratch[warp] = a;
        __syncthreads();
        if (warp == 0) {
            Stats<NSUM> p;
            if (lane < NW) p = scratch[lane];
            else stats_init(p);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                if (o < NW) {
                    Stats<NSUM> b = stats_shfl_xor<NSUM, AM, SM, MX>(p, o);
                    const bool low = (lane & o) == 0;
                    Stats<NSUM> lo = low ? p : b, hi = low ? b : p;
                    stats_merge<NSUM, AM, SM, MX>(lo, hi);
                    p = lo;
                }
            }
            if (lane == 0) scratch[NW] = p;
        }
        __syncthreads();
        a = scratch[NW];
    }
}

// ---------------------------------------------------------------------------------------------
// Gaussian patch lookup: value at (x, y) of the map whose patch is centred on (cx, cy)
//   = tab[dx^2 + dy^2] if |dx| <= tmp and |dy| <= tmp else 0     
// `tab` lives in shared memory.
// ---------------------------------------------------------------------------------------------
struct Centre {
    int x, y;  // y = INT_MIN/2 when nothing is pasted
};
constexpr int kNoPaste = -(1 << 28);

__device__ __forceinline__ float patch_at(const float* tab, int tmp, Centre c, int x, int y) {
    const int dx = x - c.x, dy = y - c.y;
    const unsigned span = 2u * static_cast<unsigned>(tmp);
    if (static_cast<unsigned>(dy + tmp) <= span && static_cast<unsigned>(dx + tmp) <= span)
        return tab[dx * dx + dy * dy];
    return 0.0f;
}
__device__ __forceinline__ float4 patch_at4(const float* tab, int tmp, Centre c, int x0, int y) {
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
    const int dy = y - c.y;
    const unsigned span = 2u * static_cast<unsigned>(tmp);
    if (static_cast<unsigned>(dy + tmp) <= span) {
        const int dx = x0 - c.x, d2 = dy * dy;
        if (static_cast<unsigned>(dx + 0 + tmp) <= span) t.x = tab[(dx + 0) * (dx + 0) + d2];
        if (static_cast<unsigned>(dx + 1 + tmp) <= span) t.y = tab[(dx + 1) * (dx + 1) + d2];
        if (static_cast<unsigned>(dx + 2 + tmp) <= span) t.z = tab[(dx + 2) * (dx + 2) + d2];
        if (static_cast<unsigned>(dx + 3 + tmp) <= span) t.w = tab[(dx + 3) * (dx + 3) + d2];
    }
    return t;
}

__device__ __forceinline__ void load_table(float* s_tab, const float* __restrict__ tab, int tmp) {
    const int n = 2 * tmp * tmp + 1;
    for (int i = threadIdx.x; i < n; i += blockDim.x) s_tab[i] = tab[i];
}
inline size_t table_bytes(int tmp) { return sizeof(float) * static_cast<size_t>(2 * tmp * tmp + 1); }

// centre of a generated target.  uda/dataset/util.py:36-46: mu = int(joint / stride + 0.5) in
// float64 with truncation toward zero; out of the map -> weight 0; pasted iff weight > 0.5.
__device__ __forceinline__ Centre target_centre(double jx, double jy, float vis, double sx, double sy, int W, int H,
                                                float& weight) {
    const double fx = trunc(__dadd_rn(__ddiv_rn(jx, sx), 0.5));
    const double fy = trunc(__dadd_rn(__ddiv_rn(jy, sy), 0.5));
    const bool inside = (fx >= 0.0) && (fx < static_cast<double>(W)) && (fy >= 0.0) && (fy < static_cast<double>(H));
    weight = inside ? vis : 0.0f;
    Centre c;
    c.x = inside ? static_cast<int>(fx) : 0;
    c.y = inside ? static_cast<int>(fy) : 0;
    if (!(inside && vis > 0.5f)) c.y = kNoPaste;  // nothing pasted: decode of the all-zero map is (0,0)
    return c;
}

// ---------------------------------------------------------------------------------------------
// decode + PCK arithmetic
// ---------------------------------------------------------------------------------------------
// utils/keypoint_detection.py:26-34 - index -> float32, x = idx % W, y = floor(idx / W), masked by max > 0
__device__ __forceinline__ void decode_xy(ArgMax a, int W, float& px, float& py) {
    const float fi = static_cast<float>(a.i), fw = static_cast<float>(W);
    const float keep = (a.v > 0.0f) ? 1.0f : 0.0f;
    px = fmodf(fi, fw) * keep;
    py = floorf(__fdiv_rn(fi, fw)) * keep;
}

// utils/keypoint_detection.py:44-47,77: valid iff tx > 1 and ty > 1; d = ||p/norm - t/norm||_2 in float64
// with norm = (H/10, W/10) applied to (x, y); numpy's norm is sqrt(ddot) = sqrt(fma(b, b, a*a)) on
// FMA hosts (OpenBLAS), which only matters at non-power-of-two sizes; hit iff d < thr.
__device__ __forceinline__ void pck_one(float px, float py, float tx, float ty, int H, int W, double thr, int& valid,
                                        int& hit) {
    valid = (tx > 1.0f && ty > 1.0f) ? 1 : 0;
    const double nx = __ddiv_rn(static_cast<double>(H), 10.0), ny = __ddiv_rn(static_cast<double>(W), 10.0);
    const double a = __dsub_rn(__ddiv_rn(static_cast<double>(px), nx), __ddiv_rn(static_cast<double>(tx), nx));
    const double b = __dsub_rn(__ddiv_rn(static_cast<double>(py), ny), __ddiv_rn(static_cast<double>(ty), ny));
    const double d = __dsqrt_rn(__fma_rn(b, b, __dmul_rn(a, a)));
    hit = (valid && d < thr) ? 1 : 0;
}

// acc[k] = hits/valid or -1; avg over acc >= 0 in joint order; cnt.  keypoint_detection.py:53-60, 80-90
__device__ __forceinline__ void pck_finalize_serial(const int* hits, const int* valid, int K, double* acc_out) {
    double total = 0.0;
    int cnt = 0;
    for (int k = 0; k < K; ++k) {
        double a = -1.0;
        if (valid[k] > 0) a = __ddiv_rn(static_cast<double>(hits[k]) * 1.0, static_cast<double>(valid[k]));
        acc_out[k] = a;
        if (a >= 0.0) {
            total = __dadd_rn(total, a);
            ++cnt;
        }
    }
    acc_out[K] = cnt != 0 ? __ddiv_rn(total, static_cast<double>(cnt)) : 0.0;
    acc_out[K + 1] = static_cast<double>(cnt);
}

__device__ __forceinline__ float f4_get(const float4& v, int c) {
    return c == 0 ? v.x : (c == 1 ? v.y : (c == 2 ? v.z : v.w));
}

// ---------------------------------------------------------------------------------------------
// per-tile KL / softmax accumulation shared by hp_loss.cu, hp_regdisp.cu and hp_pipeline.cu
// ---------------------------------------------------------------------------------------------
// online softmax update of (m, s) with the NV float4 of one tile (invalid lanes hold -inf)
template <int NV>
__device__ __forceinline__ void softmax_tile(float& m, float& s, const float4 (&p)[NV]) {
    float tm = -INFINITY;
#pragma unroll
    for (int j = 0; j < NV; ++j) tm = fmaxf(tm, fmaxf(fmaxf(p[j].x, p[j].y), fmaxf(p[j].z, p[j].w)));
    const float mn = fmaxf(m, tm);
    const float ms = (mn == -INFINITY) ? 0.0f : mn;  // all -inf so far: keep exp2(-inf) = 0, not NaN
    const float mb = -ms * kLog2e;
    float acc = s * exp_diff(m, ms);
#pragma unroll
    for (int j = 0; j < NV; ++j) {  // single-instruction MUFU.EX2 (2^-22 relative error, far inside the 1e-5 bar)
        acc += ex2_approx(fmaf(p[j].x, kLog2e, mb)) + ex2_approx(fmaf(p[j].y, kLog2e, mb));
        acc += ex2_approx(fmaf(p[j].z, kLog2e, mb)) + ex2_approx(fmaf(p[j].w, kLog2e, mb));
    }
    s = acc;
    m = mn;
}

// KL target sums for one element: sum[0] += u, sum[1] += u*p, sum[2] += xlogy(u, u) / ln 2 (the caller scales by ln 2).
// Branch-free: at u == 0 the clamp keeps lg2 finite, so the term is exactly 0 (xlogy); u < 0 gives lg2(tiny) * u, a
// finite value where torch gives NaN - a negative (target + epsilon) is outside every caller's domain (targets are
// clamped heatmaps); a NaN u propagates through the multiplication.
__device__ __forceinline__ void kl_elem(float (&sum)[3], float p, float u) {
    sum[0] += u;
    sum[1] = fmaf(u, p, sum[1]);
    sum[2] = fmaf(u, lg2_approx(fmaxf(u, 1.17549435e-38f)), sum[2]);
}

// per-map KL value from the reduced statistics:  L = (sum u ln u - sum u p)/S - ln S + lse,  lse = m + ln(s).
// float32 logarithms (two FP64 logs on one thread cost ~1,500 cycles on B200 and sat on every block's critical
// path); the two logs are merged, ln(s / S), so their rounding does not double, and the few additions run in float64.
// (Sulg2u = sum u lg2 u as accumulated by kl_elem)
__device__ __forceinline__ double kl_finish(float m, float s, float Su, float Sup, float Sulg2u, float& lse_out) {
    const float Sulogu = Sulg2u * kLn2;
    lse_out = m + logf(s);
    return static_cast<double>((Sulogu - Sup) / Su) + static_cast<double>(m) + static_cast<double>(logf(s / Su));
}

// "last block done" election.  counter must be zero on entry; it is reset by the winner.
// Every thread that published block results fences its own writes (cheap when it wrote nothing),
// then one atomic per block.
__device__ __forceinline__ bool last_block_arrives(unsigned int* counter, unsigned int n_blocks) {
    __shared__ bool s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int prev = atomicAdd(counter, 1u);
        s_last = (prev == n_blocks - 1);
        if (s_last) __threadfence();
    }
    __syncthreads();
    return s_last;
}

// ---- reductions over per-map values written by OTHER blocks (read after the last-block ticket) -----------------
// L2 loads (ld.global.cg; the writers fenced before their tickets), issued eight at a time so that their
// latencies overlap: a dependent chain of ~40 single volatile loads per thread costs tens of microseconds.
// Summation order is fixed (ascending index per thread): deterministic.
__device__ __forceinline__ double strided_sum_f64(const float* x, int n, int first, int stride) {
    double acc = 0.0;
    for (int i0 = first; i0 < n; i0 += 8 * stride) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = (i0 + u * stride < n) ? __ldcg(x + i0 + u * stride) : 0.0f;
#pragma unroll
        for (int u = 0; u < 8; ++u)
            if (i0 + u * stride < n) acc += static_cast<double>(v[u]);
    }
    return acc;
}
// per_sample[b] = mean over the K joints of per_map[b*K + k]   (loss.py:158, 'none' reduction of the KL loss)
__device__ __forceinline__ void per_sample_means(const float* per_map, int B, int K, float* per_sample, int t, int nt) {
    for (int b = t; b < B; b += nt)
        per_sample[b] = static_cast<float>(strided_sum_f64(per_map + static_cast<size_t>(b) * K, K, 0, 1) / static_cast<double>(K));
}

// Exact, order-free accumulation of per-map float losses as 64-bit integers: acc[0] += round(v * 2^40), acc[4] +=
// round(residual * 2^23) (so values down to 2^-63 still count exactly), acc[1..3] count NaN / +inf / -inf.
// Integer sums make the 'mean' independent of the grid and of the schedule, and the last block reads FIVE numbers
// instead of walking every per-map value.
constexpr float kFxAccLimit = 2097152.0f;  // |per-map loss| >= 2^21 counts as infinite
constexpr int kFxAccWords = 5;
__device__ __forceinline__ void fx_acc_add(unsigned long long* acc, float v) {
    if (fabsf(v) < kFxAccLimit) {
        const double d = static_cast<double>(v) * 1099511627776.0;  // exact
        const long long hi = __double2ll_rn(d);
        const long long lo = __double2ll_rn((d - static_cast<double>(hi)) * 8388608.0);
        atomicAdd(&acc[0], static_cast<unsigned long long>(hi));
        if (lo != 0) atomicAdd(&acc[4], static_cast<unsigned long long>(lo));
    } else {
        atomicAdd(&acc[(v != v) ? 1 : (v > 0.0f ? 2 : 3)], 1ull);
    }
}
__device__ __forceinline__ float fx_mean(const long long (&v)[kFxAccWords], int n) {
    if (v[1] != 0 || (v[2] != 0 && v[3] != 0)) return __int_as_float(0x7fc00000);
    if (v[2] != 0) return INFINITY;
    if (v[3] != 0) return -INFINITY;
    const double total = ldexp(static_cast<double>(v[0]), -40) + ldexp(static_cast<double>(v[4]), -63);
    return static_cast<float>(total / static_cast<double>(n));
}
// last block, one thread: read the workspace accumulators, restore their zero state, return the mean
__device__ __forceinline__ float fx_mean_from_workspace(unsigned long long* acc, int n) {
    long long v[kFxAccWords];
    for (int i = 0; i < kFxAccWords; ++i) {
        v[i] = static_cast<long long>(*reinterpret_cast<volatile unsigned long long*>(&acc[i]));
        acc[i] = 0ull;
    }
    return fx_mean(v, n);
}

// the same election when only a few threads published block results: only the writers fence (a __threadfence by
// every thread of every block showed up as the 'membar' stall of the loss kernels)
__device__ __forceinline__ bool last_block_arrives_writers(unsigned int* counter, unsigned int n_blocks, bool i_wrote) {
    __shared__ bool s_last1;
    if (i_wrote) __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int prev = atomicAdd(counter, 1u);
        s_last1 = (prev == n_blocks - 1);
        if (s_last1) __threadfence();
    }
    __syncthreads();
    return s_last1;
}

// deterministic block-wide float64 sum of a float array (fixed shape tree, independent of timing)
template <int NT>
__device__ __forceinline__ double block_sum_f32_as_f64(const float* __restrict__ x, int n, double* s_buf) {
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += NT) acc += static_cast<double>(x[i]);
    s_buf[threadIdx.x] = acc;
    __syncthreads();
#pragma unroll
    for (int o = NT / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) s_buf[threadIdx.x] += s_buf[threadIdx.x + o];
        __syncthreads();
    }
    const double r = s_buf[0];
    __syncthreads();
    return r;
}

// workspace layout (all ops): block counter, spare, int32 hits/valid counters, then eight 64-bit
// accumulators (the fused pipeline's fixed-point loss sums and non-finite counters)
struct Workspace {
    unsigned int counter;
    unsigned int spare;
    int counts[2 * HP_MAX_K];
    unsigned long long acc[8];
};

}  // namespace hp
