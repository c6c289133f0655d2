// hp_bilinear_block.cuh - nn.Upsample(mode='bilinear', align_corners=False) at the EXACT x2 / x4 scales as a static tap
// pattern over 4x4 output blocks, from a source map staged in shared memory (train1.py:410-424: 16 -> 64, 32 -> 64, 16 -> 32).
// Shared by the fusion kernels (hp_fusion_block.cuh, which documents the pattern and checks it against make_tap) and by the
// dense disparity kernel when it builds the fused map itself (hp_regdisp_dense.cuh, hp_regdisp_fwd_heads).
// Arithmetic and operation order: fmul + ffma horizontally, fmul + ffma vertically (bit-identical to the row-walking kernels).
#pragma once
#include <cstdint>

namespace hp {

__device__ __forceinline__ float lds_f32(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ float2 lds_f32x2(uint32_t addr) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
    return v;
}

template <int S>
struct BlockAxis {           // one axis of one source as seen by a lane / a block row
    float2 l0a, l1a;         // weights of outputs 0, 1 (first-block clamp folded in)
    float2 l0b, l1b;         // weights of outputs 2, 3
    bool first;
};
template <int S>
__device__ __forceinline__ BlockAxis<S> block_axis(int m) {
    BlockAxis<S> ax;
    ax.first = (m == 0);
    if (S == 2) {
        ax.l1a = make_float2(ax.first ? 0.0f : 0.75f, 0.25f);
        ax.l1b = make_float2(0.75f, 0.25f);
    } else {
        ax.l1a = ax.first ? make_float2(0.0f, 0.0f) : make_float2(0.625f, 0.875f);
        ax.l1b = make_float2(0.125f, 0.375f);
    }
    ax.l0a = make_float2(1.0f - ax.l1a.x, 1.0f - ax.l1a.y);
    ax.l0b = make_float2(1.0f - ax.l1b.x, 1.0f - ax.l1b.y);
    return ax;
}
// byte offsets of a lane's distinct source columns inside a source row
template <int S>
struct BlockCols {
    uint32_t a, b, d;  // S=2: A, (B,C) as one 8-byte load, D;  S=4: A, B, C (d)
};
template <int S>
__device__ __forceinline__ BlockCols<S> block_cols(int n, int w) {
    BlockCols<S> c;
    if (S == 2) {
        c.a = 4u * static_cast<uint32_t>(max(2 * n - 1, 0));
        c.b = 8u * static_cast<uint32_t>(n);
        c.d = 4u * static_cast<uint32_t>(min(2 * n + 2, w - 1));
    } else {
        c.a = 4u * static_cast<uint32_t>(max(n - 1, 0));
        c.b = 4u * static_cast<uint32_t>(n);
        c.d = 4u * static_cast<uint32_t>(min(n + 1, w - 1));
    }
    return c;
}
// one source row (shared-memory byte address) interpolated to the lane's four output columns
template <int S>
__device__ __forceinline__ void block_hrow(uint32_t row, const BlockCols<S>& c, const BlockAxis<S>& ax, float2 (&t)[2]) {
    float2 a01, b01, a23, b23;
    if (S == 2) {
        const float vA = lds_f32(row + c.a);
        const float2 vBC = lds_f32x2(row + c.b);
        const float vD = lds_f32(row + c.d);
        a01 = make_float2(vA, vBC.x);
        b01 = make_float2(ax.first ? vBC.y : vBC.x, vBC.y);
        a23 = make_float2(vBC.x, vBC.y);
        b23 = make_float2(vBC.y, vD);
    } else {
        const float vA = lds_f32(row + c.a), vB = lds_f32(row + c.b), vC = lds_f32(row + c.d);
        const float dup = ax.first ? vC : vB;
        a01 = make_float2(vA, vA);
        b01 = make_float2(dup, dup);
        a23 = make_float2(vB, vB);
        b23 = make_float2(vC, vC);
    }
    t[0] = __ffma2_rn(ax.l1a, b01, __fmul2_rn(ax.l0a, a01));
    t[1] = __ffma2_rn(ax.l1b, b23, __fmul2_rn(ax.l0b, a23));
}

// the interpolated source rows a lane carries down its column of blocks, and the vertical blend of a block
template <int S>
struct BlockRows {
    float2 A[2], B[2], C[2], D[2];  // D unused for S=4
};
template <int S>
__device__ __forceinline__ void block_rows_start(BlockRows<S>& R, uint32_t base, int w, int in_h, int m, const BlockCols<S>& c,
                                                 const BlockAxis<S>& ax) {
    const uint32_t stride = 4u * static_cast<uint32_t>(w);
    if (S == 2) {
        block_hrow<S>(base + stride * static_cast<uint32_t>(max(2 * m - 1, 0)), c, ax, R.A);
        block_hrow<S>(base + stride * static_cast<uint32_t>(2 * m), c, ax, R.B);
    } else {
        block_hrow<S>(base + stride * static_cast<uint32_t>(max(m - 1, 0)), c, ax, R.A);
        block_hrow<S>(base + stride * static_cast<uint32_t>(m), c, ax, R.B);
    }
    (void)in_h;
}
// block m: load the new rows, produce the four blended output rows v[r][0..1] (columns (0,1), (2,3))
template <int S>
__device__ __forceinline__ void block_rows_blend(BlockRows<S>& R, uint32_t base, int w, int in_h, int m, const BlockCols<S>& c,
                                                 const BlockAxis<S>& ax, float2 (&v)[4][2]) {
    const uint32_t stride = 4u * static_cast<uint32_t>(w);
    const bool first = (m == 0);
    if (S == 2) {
        block_hrow<S>(base + stride * static_cast<uint32_t>(2 * m + 1), c, ax, R.C);
        block_hrow<S>(base + stride * static_cast<uint32_t>(min(2 * m + 2, in_h - 1)), c, ax, R.D);
        const float l1r0 = first ? 0.0f : 0.75f, l0r0 = 1.0f - l1r0;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const float2 dup = first ? R.C[h] : R.B[h];
            v[0][h] = __ffma2_rn(make_float2(l1r0, l1r0), dup, __fmul2_rn(make_float2(l0r0, l0r0), R.A[h]));
            v[1][h] = __ffma2_rn(make_float2(0.25f, 0.25f), R.C[h], __fmul2_rn(make_float2(0.75f, 0.75f), R.B[h]));
            v[2][h] = __ffma2_rn(make_float2(0.75f, 0.75f), R.C[h], __fmul2_rn(make_float2(0.25f, 0.25f), R.B[h]));
            v[3][h] = __ffma2_rn(make_float2(0.25f, 0.25f), R.D[h], __fmul2_rn(make_float2(0.75f, 0.75f), R.C[h]));
            R.A[h] = R.C[h];  // carried into block m + 1
            R.B[h] = R.D[h];
        }
    } else {
        block_hrow<S>(base + stride * static_cast<uint32_t>(min(m + 1, in_h - 1)), c, ax, R.C);
        const float l1r0 = first ? 0.0f : 0.625f, l1r1 = first ? 0.0f : 0.875f;
        const float l0r0 = 1.0f - l1r0, l0r1 = 1.0f - l1r1;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const float2 dup = first ? R.C[h] : R.B[h];
            v[0][h] = __ffma2_rn(make_float2(l1r0, l1r0), dup, __fmul2_rn(make_float2(l0r0, l0r0), R.A[h]));
            v[1][h] = __ffma2_rn(make_float2(l1r1, l1r1), dup, __fmul2_rn(make_float2(l0r1, l0r1), R.A[h]));
            v[2][h] = __ffma2_rn(make_float2(0.125f, 0.125f), R.C[h], __fmul2_rn(make_float2(0.875f, 0.875f), R.B[h]));
            v[3][h] = __ffma2_rn(make_float2(0.375f, 0.375f), R.C[h], __fmul2_rn(make_float2(0.625f, 0.625f), R.B[h]));
            R.A[h] = R.B[h];
            R.B[h] = R.C[h];
        }
    }
}

}  // namespace hp
