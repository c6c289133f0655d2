// hp_fusion_block.cuh - multiscale fusion for the EXACT x2 / x4 scales (train1.py:410-424: 16 -> 64, 32 -> 64, 16 -> 32;
// BASELINE.json configs[3]: 32 / 64 / 128), the production shape of hp_fuse_multiscale and hp_fuse_decode_pck.
//
// nn.Upsample(mode='bilinear') is align_corners=False: src = scale*(dst+0.5)-0.5 clamped at 0.  At an integer factor S
// the taps of four consecutive outputs 4m..4m+3 are a STATIC pattern over three (S=4) or four (S=2) consecutive source
// positions with constant weights; only the first block of an axis (clamp at 0: weights (1, 0), taps (0, 1)) and the
// last one (second tap clamped to in-1) differ, and both fit the pattern with one select and two clamped offsets:
//   S=2: sources A=max(2m-1,0) B=2m C=2m+1 D=min(2m+2,in-1);  outputs (A,B') (B,C) (B,C) (C,D), l1 = .75 .25 .75 .25
//   S=4: sources A=max(m-1,0)  B=m  C=min(m+1,in-1);          outputs (A,B') (A,B') (B,C) (B,C), l1 = .625 .875 .125 .375
//   B' = B, except in the first block where the clamp makes the taps (0, 1): B' = C (S=2: index 1 = 2m+1; S=4: m+1)
//   and the weights of the clamped outputs (1, 0).
// So a lane owns a column of 4x4 output blocks: per block it interpolates the two / one NEW source rows horizontally
// (the others are carried in registers from the block above), blends vertically with immediate weights and streams
// its four float4 of `hi` - no cached-row bookkeeping, no tap table, ~40 instructions per output float4 instead of
// ~150 (the row-walking kernels in hp_fusion.cu, which remain the path for every other geometry and serve the
// NaN rescan here).  Arithmetic and operation order are those of fused_row4 (fmul + ffma horizontally, fmul + ffma
// vertically, fmul / ffma / ffma over the sources), so the results are bit-identical to the row-walking kernels.
// The low-resolution maps are staged in shared memory by the copy engine, double-buffered per block, exactly as in
// fuse_staged_kernel (ticket hand-over, no block barrier in the loop).
#pragma once

namespace hp {

// debug guard: the pattern above must reproduce make_tap for outputs 4m..4m+3 (exact in fp32 for S in {2, 4})
template <int S>
__device__ __forceinline__ bool block_axis_matches(int m, int in_size) {
    const BlockAxis<S> ax = block_axis<S>(m);
    const float l1[4] = {ax.l1a.x, ax.l1a.y, ax.l1b.x, ax.l1b.y};
    bool ok = true;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const Tap t = make_tap(4 * m + c, 1.0f / static_cast<float>(S), in_size, in_size * S);
        int i0, i1;
        if (S == 2) {
            const int A = max(2 * m - 1, 0), B = 2 * m, Cc = 2 * m + 1, D = min(2 * m + 2, in_size - 1);
            i0 = c == 0 ? A : (c == 3 ? Cc : B);
            i1 = c == 0 ? (m == 0 ? Cc : B) : (c == 3 ? D : Cc);
        } else {
            const int A = max(m - 1, 0), B = m, Cc = min(m + 1, in_size - 1);
            i0 = c < 2 ? A : B;
            i1 = c < 2 ? (m == 0 ? Cc : B) : Cc;
        }
        ok = ok && t.i0 == i0 && t.i1 == i1 && t.l1 == l1[c] && t.l0 == 1.0f - l1[c];
    }
    return ok;
}

struct BlockWalk {
    int cb;     // 4-column blocks per output row (W / 4)
    int rps;    // block rows walked side by side by one warp (32 / cb)
    int strip;  // consecutive block rows per lane
};

// the whole-map NaN rescan of the closing warp (rare): the row-walking code with numpy's exact argmax rules
static __device__ __noinline__ ArgMax block_nan_rescan(const FuseSrc& f, uint32_t lo_s, uint32_t mid_s, const RowTap* s_rows,
                                                       const float* hi_map, int lane) {
    const int cpr = f.W / 4, rps = 32 / cpr;
    const int x0 = (lane % cpr) * 4, ro = lane / cpr;
    RowSource lo, mid;
    row_source_init(lo, x0, f.sx_lo, f.wl, f.W);
    row_source_init(mid, x0, f.mid ? f.sx_mid : 1.0f, f.mid ? f.wm : f.W, f.W);
    lo.base_s = lo_s;
    mid.base_s = mid_s;
    ArgMax sx = am_init();
    for (int r = ro; r < f.H; r += rps) {
        const float4 hb = hi_map ? ldg_stream4(reinterpret_cast<const float4*>(hi_map + r * f.W + x0)) : make_float4(0.f, 0.f, 0.f, 0.f);
        am_scan4<true>(sx, fused_row4<true>(f, lo, mid, s_rows, r, hb), r * f.W + x0);
    }
    return warp_argmax_rows(sx, lane);
}

constexpr int kBlockMaxWarps = 4;
struct BlockCtl {
    unsigned long long full[2];  // mbarriers: sources of buffer b have landed
    int done[2];                 // warps that have finished the map in buffer b
    int nan[2];
    int bad_geometry;
    ArgMax am[2][kBlockMaxWarps];
};

// SL: scale of `lo` (2 or 4);  SM: scale of `mid` (2, or 0 = no mid source);  STRIP: block rows per lane at compile
// time (0 = g.strip at run time) - with a known even trip count the loop is unrolled by two and the rows carried from
// one block to the next (and the prefetched `hi` rows) are renamed instead of moved (ncu attributed 9 % of all
// executed instructions to those loop-edge moves).  Measured on B200: strips of 2 (64x64 outputs) gain 6 %
// (66.4 -> 62.3 us); strips of 8 (128x128) spill at the 128-register bound of four blocks per SM and lose 6 %
// (three blocks per SM with 164 registers: 109.8 vs 104.6 us - the kernel needs its 16 warps), so they keep the run-time loop.
// PAIR (train1.py:410-424 in ONE launch): besides `out` = a_lo up4(lo) + a_mid up2(mid) the block also writes
// out2 = a2 * up2(lo) at half the size (`target0 = up32(y_adv3)` next to `target5`): 64 of its threads take one 4x4 block of
// the second map each, from the same staged source - the second launch and its re-read of `lo` disappear.
template <int SL, int SM, bool DECODE, int STRIP, bool PAIR = false>
__global__ void __launch_bounds__(32 * kBlockMaxWarps, 4)
    fuse_block_kernel(const FuseSrc f, const BlockWalk g, int n_maps, float* __restrict__ out, const float* __restrict__ tgt_xy,
                      int K, double thr, float* __restrict__ pred_xy, float* __restrict__ maxvals,
                      int32_t* __restrict__ counts_out, double* __restrict__ acc_out, Workspace* __restrict__ ws,
                      float* __restrict__ out2 = nullptr, float a2 = 0.0f) {
    extern __shared__ __align__(128) unsigned char s_raw[];
    __shared__ BlockCtl ctl;
    constexpr int SMX = SM == 0 ? 2 : SM;  // (template argument of the unused mid helpers)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n_warps = blockDim.x >> 5;
    const int HW = f.H * f.W;
    const int lo_elems = f.hl * f.wl, mid_elems = SM ? f.hm * f.wm : 0;
    const uint32_t lo_bytes = 4u * lo_elems, mid_bytes = 4u * mid_elems, buf_bytes = lo_bytes + mid_bytes;
    RowTap* s_rows = reinterpret_cast<RowTap*>(s_raw + 2 * static_cast<size_t>(buf_bytes));  // NaN rescan only
    const int n_local = (n_maps - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
    const uint32_t full_u32 = smem_addr(&ctl.full[0]), src_u32 = smem_addr(s_raw);
    const uint64_t pol = l2_evict_first_policy();

    auto request = [&](int j) {  // one thread: sources of the block's j-th map -> buffer j & 1
        const int b = j & 1;
        const size_t map = static_cast<size_t>(blockIdx.x) + static_cast<size_t>(j) * gridDim.x;
        mbar_arrive_expect_tx(full_u32 + 8 * b, buf_bytes);
        bulk_load(src_u32 + b * buf_bytes, f.lo + map * lo_elems, lo_bytes, full_u32 + 8 * b, pol);
        if (SM) bulk_load(src_u32 + b * buf_bytes + lo_bytes, f.mid + map * mid_elems, mid_bytes, full_u32 + 8 * b, pol);
    };
    if (threadIdx.x == 0) {
        mbar_init(full_u32, 1);
        mbar_init(full_u32 + 8, 1);
        mbar_init_fence();
        ctl.done[0] = ctl.done[1] = 0;
        ctl.nan[0] = ctl.nan[1] = 0;
        ctl.bad_geometry = 0;
        if (n_local > 0) request(0);
        if (n_local > 1) request(1);
    }
    if (DECODE) {
        for (int r = threadIdx.x; r < f.H; r += blockDim.x) {
            const Tap tl = make_tap(r, f.sy_lo, f.hl, f.H);
            s_rows[r] = RowTap{tl.i0, tl.i1, tl.l0, tl.l1};
            if (SM) {
                const Tap tm = make_tap(r, f.sy_mid, f.hm, f.H);
                s_rows[f.H + r] = RowTap{tm.i0, tm.i1, tm.l0, tm.l1};
            }
        }
    }
    // this lane: column block n, block rows [m_begin, m_end)
    const int n = lane % g.cb, sub = lane / g.cb;
    const int strip = STRIP ? STRIP : g.strip;
    const int m_begin = (warp * g.rps + sub) * strip, m_end = m_begin + strip;
    constexpr int kUnroll = (STRIP >= 2 && STRIP % 2 == 0) ? 2 : 1;
    const int x0 = 4 * n;
    const BlockAxis<SL> lo_x = block_axis<SL>(n);
    const BlockCols<SL> lo_c = block_cols<SL>(n, f.wl);
    const BlockAxis<SMX> mid_x = block_axis<SMX>(n);
    const BlockCols<SMX> mid_c = block_cols<SMX>(n, SM ? f.wm : 2);
    // PAIR: thread i < (H/8) * (W/8) owns block (m2, n2) of the half-size map
    const int cb2 = f.W / 8, nb2 = (f.H / 8) * cb2;
    const int n2 = PAIR ? static_cast<int>(threadIdx.x) % cb2 : 0, m2 = PAIR ? static_cast<int>(threadIdx.x) / cb2 : 0;
    const BlockAxis<2> ax2 = block_axis<2>(n2);
    const BlockCols<2> col2 = block_cols<2>(n2, f.wl);
    __syncthreads();  // the only block barrier: control block (and tap table) are set up
    {
        bool ok = block_axis_matches<SL>(n, f.wl);
        if (SM) ok = ok && block_axis_matches<SMX>(n, f.wm);
        if (PAIR && static_cast<int>(threadIdx.x) < nb2) ok = ok && block_axis_matches<2>(n2, f.wl) && block_axis_matches<2>(m2, f.hl);
        for (int m = m_begin; m < m_end; ++m) {
            ok = ok && block_axis_matches<SL>(m, f.hl);
            if (SM) ok = ok && block_axis_matches<SMX>(m, f.hm);
        }
        if (!ok) __trap();  // the host only selects this kernel for exact scales; never silently wrong
    }

    const float2 al = make_float2(f.a_lo, f.a_lo), am2 = make_float2(f.a_mid, f.a_mid), ah = make_float2(f.a_hi, f.a_hi);
    for (int j = 0; j < n_local; ++j) {
        const int b = j & 1;
        const int map = static_cast<int>(blockIdx.x) + j * static_cast<int>(gridDim.x);
        const uint32_t lo_s = src_u32 + b * buf_bytes, mid_s = lo_s + lo_bytes;
        const float* hi = f.hi ? f.hi + static_cast<size_t>(map) * HW + x0 : nullptr;
        float* o = DECODE ? nullptr : out + static_cast<size_t>(map) * HW + x0;
        // running maximum of this lane, branch-free: the row of the first (strict) improvement and that row's four
        // values are kept, the component is resolved once after the loop (the earlier, lower-index element keeps ties)
        float best = -INFINITY;
        int best_row = 4 * m_begin;  // (an all -inf map decodes to its first element)
        float4 best_v = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
        float2 witness = make_float2(0.f, 0.f);
        float4 h[4];
#pragma unroll
        for (int u = 0; u < 4; ++u)  // the first rows of the HBM stream are requested before the sources are awaited
            h[u] = hi ? ldg_stream4(reinterpret_cast<const float4*>(hi + (4 * m_begin + u) * f.W)) : make_float4(0.f, 0.f, 0.f, 0.f);
        mbar_wait(full_u32 + 8 * b, static_cast<uint32_t>(j >> 1) & 1u);
        BlockRows<SL> RL;
        BlockRows<SMX> RM;
        block_rows_start<SL>(RL, lo_s, f.wl, f.hl, m_begin, lo_c, lo_x);
        if (SM) block_rows_start<SMX>(RM, mid_s, f.wm, f.hm, m_begin, mid_c, mid_x);
#pragma unroll kUnroll
        for (int i = 0; i < strip; ++i) {
            const int m = m_begin + i;
            float4 hn[4];
#pragma unroll
            for (int u = 0; u < 4; ++u)  // the next block's rows in flight while this block is blended
                hn[u] = (hi && i + 1 < strip) ? ldg_stream4(reinterpret_cast<const float4*>(hi + (4 * (m + 1) + u) * f.W))
                                             : make_float4(0.f, 0.f, 0.f, 0.f);
            float2 vl[4][2], vm[4][2];
            block_rows_blend<SL>(RL, lo_s, f.wl, f.hl, m, lo_c, lo_x, vl);
            if (SM) block_rows_blend<SMX>(RM, mid_s, f.wm, f.hm, m, mid_c, mid_x, vm);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                float2 r0 = __fmul2_rn(al, vl[u][0]), r1 = __fmul2_rn(al, vl[u][1]);
                if (SM) {
                    r0 = __ffma2_rn(am2, vm[u][0], r0);
                    r1 = __ffma2_rn(am2, vm[u][1], r1);
                }
                if (f.hi) {
                    r0 = __ffma2_rn(ah, make_float2(h[u].x, h[u].y), r0);
                    r1 = __ffma2_rn(ah, make_float2(h[u].z, h[u].w), r1);
                }
                const int row = 4 * m + u;
                if (DECODE) {
                    const float m4 = fmaxf(fmaxf(r0.x, r0.y), fmaxf(r1.x, r1.y));
                    const bool up = m4 > best;  // strict; a NaN row never improves (the witness sends the map to the rescan)
                    best = up ? m4 : best;
                    best_row = up ? row : best_row;
                    best_v.x = up ? r0.x : best_v.x;
                    best_v.y = up ? r0.y : best_v.y;
                    best_v.z = up ? r1.x : best_v.z;
                    best_v.w = up ? r1.y : best_v.w;
                    witness = __fadd2_rn(witness, __fadd2_rn(r0, r1));  // NaN / inf-inf witness
                } else {
                    stg_stream4(reinterpret_cast<float4*>(o + row * f.W), make_float4(r0.x, r0.y, r1.x, r1.y));
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) h[u] = hn[u];
        }
        if (PAIR && static_cast<int>(threadIdx.x) < nb2) {  // the half-size map from the same staged `lo` (still valid: before the ticket)
            BlockRows<2> R2;
            block_rows_start<2>(R2, lo_s, f.wl, f.hl, m2, col2, ax2);
            float2 v2[4][2];
            block_rows_blend<2>(R2, lo_s, f.wl, f.hl, m2, col2, ax2, v2);
            const float2 aa = make_float2(a2, a2);
            float* o2 = out2 + static_cast<size_t>(map) * (HW / 4) + (4 * m2) * (f.W / 2) + 4 * n2;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float2 r0 = __fmul2_rn(aa, v2[u][0]), r1 = __fmul2_rn(aa, v2[u][1]);
                stg_stream4(reinterpret_cast<float4*>(o2 + u * (f.W / 2)), make_float4(r0.x, r0.y, r1.x, r1.y));
            }
        }
        ArgMax am = am_init();
        bool bad = false;
        if (DECODE) {
            const int comp = (best_v.x == best) ? 0 : ((best_v.y == best) ? 1 : ((best_v.z == best) ? 2 : 3));
            am.v = best;
            am.i = best_row * f.W + x0 + comp;
            am = warp_argmax_rows(am, lane);
            const float w = witness.x + witness.y;
            bad = __any_sync(0xffffffffu, w != w);
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
            last = (atomicAdd(&ctl.done[b], 1) == n_warps - 1) ? 1 : 0;
            if (last) __threadfence_block();
        }
        last = __shfl_sync(0xffffffffu, last, 0);
        if (last) {
            if (DECODE) {
                ArgMax a = ctl.am[b][0];
                for (int w = 1; w < n_warps; ++w) a = am_merge(a, ctl.am[b][w]);
                if (*reinterpret_cast<volatile int*>(&ctl.nan[b]))
                    a = block_nan_rescan(f, lo_s, mid_s, s_rows, f.hi ? f.hi + static_cast<size_t>(map) * HW : nullptr, lane);
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

// ---------------------------------------------------------------------------------------------------------------------
// The same walk with the HIGH-RESOLUTION map staged by the copy engine too (configs[3]: 32 / 64 / 128, fuse + decode +
// PCK).  In fuse_block_kernel a lane streams its four float4 of `hi` per block row with plain loads, one block row
// ahead: ncu (profiles/r1_final_kernels_summary.txt) shows the warps waiting on exactly those loads (long-scoreboard
// 2.1 per issued instruction at 57 % issue utilisation) and 16 register moves per block row that hand the prefetched
// rows on.  Here a block is 8 warps (2 blocks per SM, the same 16 warps), a warp owns 16 output rows = 8 KB of `hi`,
// split into two 4 KB halves with their own mbarriers: the half a warp has finished is re-requested for the block's
// NEXT map at once, so each half has half a map of lead time and the blend never waits on global memory.
constexpr int kBlockHsWarps = 8;
struct BlockHsCtl {
    unsigned long long full[2];               // lo / mid sources of buffer b have landed
    unsigned long long hfull[kBlockHsWarps][4];  // part h of warp w's rows of `hi` has landed
    int done[2];
    int nan[2];
    ArgMax am[2][kBlockHsWarps];
};

#ifndef HP_FUSE_HS_SUB
#define HP_FUSE_HS_SUB 2  // (1: four 2 KB parts per warp measured slower - 98.5 vs 94.4 us per 256 samples)
#endif
template <int SL, int SM, bool DECODE>
__global__ void __launch_bounds__(32 * kBlockHsWarps, 2)
    fuse_block_hs_kernel(const FuseSrc f, int n_maps, float* __restrict__ out, const float* __restrict__ tgt_xy, int K, double thr,
                         float* __restrict__ pred_xy, float* __restrict__ maxvals, int32_t* __restrict__ counts_out,
                         double* __restrict__ acc_out, Workspace* __restrict__ ws, const PeerLink link, const int defer,
                         long long* __restrict__ partial_out, double* __restrict__ result_out) {
    extern __shared__ __align__(128) unsigned char s_raw[];
    __shared__ BlockHsCtl ctl;
    constexpr int SMX = SM == 0 ? 2 : SM;
    constexpr int STRIP = 4, SUB = HP_FUSE_HS_SUB, NSUB = STRIP / SUB;  // block rows per lane, per separately requested part
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int HW = f.H * f.W;
    const int lo_elems = f.hl * f.wl, mid_elems = SM ? f.hm * f.wm : 0;
    const uint32_t lo_bytes = 4u * lo_elems, mid_bytes = 4u * mid_elems, buf_bytes = lo_bytes + mid_bytes;
    RowTap* s_rows = reinterpret_cast<RowTap*>(s_raw + 2 * static_cast<size_t>(buf_bytes));  // NaN rescan only
    const uint32_t rows_bytes = ((static_cast<uint32_t>(sizeof(RowTap)) * 2u * f.H + 127u) / 128u) * 128u;
    const uint32_t half_bytes = static_cast<uint32_t>(SUB * 4 * f.W) * 4u;  // 8 rows of `hi`
    const int n_local = (n_maps - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
    const uint32_t full_u32 = smem_addr(&ctl.full[0]), src_u32 = smem_addr(s_raw);
    const uint32_t hbar_u32 = smem_addr(&ctl.hfull[warp][0]);
    const uint32_t hs_u32 = src_u32 + 2u * buf_bytes + rows_bytes + static_cast<uint32_t>(warp) * NSUB * half_bytes;
    const uint64_t pol = l2_evict_first_policy();
    const int m_begin = warp * STRIP;  // one block row per warp pass (W / 4 == 32 column blocks)

    auto request = [&](int j) {  // one thread: lo / mid of the block's j-th map -> buffer j & 1
        const int b = j & 1;
        const size_t map = static_cast<size_t>(blockIdx.x) + static_cast<size_t>(j) * gridDim.x;
        mbar_arrive_expect_tx(full_u32 + 8 * b, buf_bytes);
        bulk_load(src_u32 + b * buf_bytes, f.lo + map * lo_elems, lo_bytes, full_u32 + 8 * b, pol);
        if (SM) bulk_load(src_u32 + b * buf_bytes + lo_bytes, f.mid + map * mid_elems, mid_bytes, full_u32 + 8 * b, pol);
    };
    auto request_hi = [&](int j, int half) {  // lane 0 of a warp: its rows [16 warp + 8 half, + 8) of the block's j-th map
        const size_t map = static_cast<size_t>(blockIdx.x) + static_cast<size_t>(j) * gridDim.x;
        const float* src = f.hi + map * HW + static_cast<size_t>(4 * (m_begin + half * SUB)) * f.W;
        mbar_arrive_expect_tx(hbar_u32 + 8 * half, half_bytes);
        bulk_load(hs_u32 + half * half_bytes, src, half_bytes, hbar_u32 + 8 * half, pol);
    };
    if (lane == 0) {
        for (int q = 0; q < NSUB; ++q) mbar_init(hbar_u32 + 8 * q, 1);
        if (warp == 0) {
            mbar_init(full_u32, 1);
            mbar_init(full_u32 + 8, 1);
        }
        mbar_init_fence();
        if (n_local > 0)
            for (int q = 0; q < NSUB; ++q) request_hi(0, q);
        if (warp == 0) {
            ctl.done[0] = ctl.done[1] = 0;
            ctl.nan[0] = ctl.nan[1] = 0;
            if (n_local > 0) request(0);
            if (n_local > 1) request(1);
        }
    }
    if (DECODE) {
        for (int r = threadIdx.x; r < f.H; r += blockDim.x) {
            const Tap tl = make_tap(r, f.sy_lo, f.hl, f.H);
            s_rows[r] = RowTap{tl.i0, tl.i1, tl.l0, tl.l1};
            if (SM) {
                const Tap tm = make_tap(r, f.sy_mid, f.hm, f.H);
                s_rows[f.H + r] = RowTap{tm.i0, tm.i1, tm.l0, tm.l1};
            }
        }
    }
    const int n = lane, x0 = 4 * n;
    const BlockAxis<SL> lo_x = block_axis<SL>(n);
    const BlockCols<SL> lo_c = block_cols<SL>(n, f.wl);
    const BlockAxis<SMX> mid_x = block_axis<SMX>(n);
    const BlockCols<SMX> mid_c = block_cols<SMX>(n, SM ? f.wm : 2);
    __syncthreads();  // the only block barrier: control block (and tap table) are set up
    {
        bool ok = block_axis_matches<SL>(n, f.wl);
        if (SM) ok = ok && block_axis_matches<SMX>(n, f.wm);
        for (int m = m_begin; m < m_begin + STRIP; ++m) {
            ok = ok && block_axis_matches<SL>(m, f.hl);
            if (SM) ok = ok && block_axis_matches<SMX>(m, f.hm);
        }
        if (!ok) __trap();  // the host only selects this kernel for exact scales; never silently wrong
    }

    const float2 al = make_float2(f.a_lo, f.a_lo), am2 = make_float2(f.a_mid, f.a_mid), ah = make_float2(f.a_hi, f.a_hi);
    const uint32_t my_hi = hs_u32 + 16u * static_cast<uint32_t>(lane);  // this lane's float4 of a staged row
    const uint32_t row_bytes = 4u * static_cast<uint32_t>(f.W);
    for (int j = 0; j < n_local; ++j) {
        const int b = j & 1;
        const int map = static_cast<int>(blockIdx.x) + j * static_cast<int>(gridDim.x);
        const uint32_t lo_s = src_u32 + b * buf_bytes, mid_s = lo_s + lo_bytes;
        float* o = DECODE ? nullptr : out + static_cast<size_t>(map) * HW + x0;
        float best = -INFINITY;
        int best_row = 4 * m_begin;
        float4 best_v = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
        float2 witness = make_float2(0.f, 0.f);
        mbar_wait(full_u32 + 8 * b, static_cast<uint32_t>(j >> 1) & 1u);
        BlockRows<SL> RL;
        BlockRows<SMX> RM;
        block_rows_start<SL>(RL, lo_s, f.wl, f.hl, m_begin, lo_c, lo_x);
        if (SM) block_rows_start<SMX>(RM, mid_s, f.wm, f.hm, m_begin, mid_c, mid_x);
#pragma unroll
        for (int i = 0; i < STRIP; ++i) {
            const int m = m_begin + i;
            const int half = i / SUB;
            if (i % SUB == 0) mbar_wait(hbar_u32 + 8 * half, static_cast<uint32_t>(j) & 1u);
            float4 h[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint32_t addr = my_hi + half * half_bytes + static_cast<uint32_t>((i % SUB) * 4 + u) * row_bytes;
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(h[u].x), "=f"(h[u].y), "=f"(h[u].z), "=f"(h[u].w) : "r"(addr));
            }
            if (i % SUB == SUB - 1) {  // this half has been read out: the block's next map takes its place
                __syncwarp();
                if (lane == 0 && j + 1 < n_local) request_hi(j + 1, half);
            }
            float2 vl[4][2], vm[4][2];
            block_rows_blend<SL>(RL, lo_s, f.wl, f.hl, m, lo_c, lo_x, vl);
            if (SM) block_rows_blend<SMX>(RM, mid_s, f.wm, f.hm, m, mid_c, mid_x, vm);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                float2 r0 = __fmul2_rn(al, vl[u][0]), r1 = __fmul2_rn(al, vl[u][1]);
                if (SM) {
                    r0 = __ffma2_rn(am2, vm[u][0], r0);
                    r1 = __ffma2_rn(am2, vm[u][1], r1);
                }
                r0 = __ffma2_rn(ah, make_float2(h[u].x, h[u].y), r0);
                r1 = __ffma2_rn(ah, make_float2(h[u].z, h[u].w), r1);
                const int row = 4 * m + u;
                if (DECODE) {
                    const float m4 = fmaxf(fmaxf(r0.x, r0.y), fmaxf(r1.x, r1.y));
                    const bool up = m4 > best;  // strict; a NaN row never improves (the witness sends the map to the rescan)
                    best = up ? m4 : best;
                    best_row = up ? row : best_row;
                    best_v.x = up ? r0.x : best_v.x;
                    best_v.y = up ? r0.y : best_v.y;
                    best_v.z = up ? r1.x : best_v.z;
                    best_v.w = up ? r1.y : best_v.w;
                    witness = __fadd2_rn(witness, __fadd2_rn(r0, r1));  // NaN / inf-inf witness
                } else {
                    stg_stream4(reinterpret_cast<float4*>(o + row * f.W), make_float4(r0.x, r0.y, r1.x, r1.y));
                }
            }
        }
        ArgMax am = am_init();
        bool bad = false;
        if (DECODE) {
            const int comp = (best_v.x == best) ? 0 : ((best_v.y == best) ? 1 : ((best_v.z == best) ? 2 : 3));
            am.v = best;
            am.i = best_row * f.W + x0 + comp;
            am = warp_argmax_rows(am, lane);
            const float w = witness.x + witness.y;
            bad = __any_sync(0xffffffffu, w != w);
        }
        // ---- ticket: the last warp of this map closes it and re-fills the lo / mid buffer ------------------------
        __syncwarp();
        int last = 0;
        if (lane == 0) {
            if (DECODE) {
                ctl.am[b][warp] = am;
                if (bad) atomicOr(&ctl.nan[b], 1);
            }
            __threadfence_block();
            last = (atomicAdd(&ctl.done[b], 1) == kBlockHsWarps - 1) ? 1 : 0;
            if (last) __threadfence_block();
        }
        last = __shfl_sync(0xffffffffu, last, 0);
        if (last) {
            if (DECODE) {
                ArgMax a = ctl.am[b][0];
                for (int w = 1; w < kBlockHsWarps; ++w) a = am_merge(a, ctl.am[b][w]);
                if (*reinterpret_cast<volatile int*>(&ctl.nan[b]))
                    a = block_nan_rescan(f, lo_s, mid_s, s_rows, f.hi + static_cast<size_t>(map) * HW, lane);
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
        if (last_block_arrives(&ws->counter, gridDim.x)) {
            if (link.world > 1) {
                // sharded (configs[3] on N GPUs): the last block sums the integer counts over the ranks itself - one warp, over
                // the NVLink peer mailboxes (hp_peer_step.cuh) - instead of a second launch; scratch = the dead source buffers
                long long* s_total = reinterpret_cast<long long*>(s_raw);
                long long* s_scratch = s_total + 144;
                double* s_acc = reinterpret_cast<double*>(s_scratch + kPeerScratchWords);
                double* s_result = s_acc + HP_MAX_K;
                if (warp == 0) {
                    const int n = 4 + 2 * K + 6;
                    for (int w = lane; w < n; w += 32) {
                        long long v = 0;
                        if (w >= 4 && w < 4 + 2 * K) {
                            v = *reinterpret_cast<volatile int*>(&ws->counts[w - 4]);
                            ws->counts[w - 4] = 0;
                        }
                        s_total[w] = v;
                    }
                    __syncwarp();
                    // defer != 0 (a step of a train): only SEND; the totals of this step land in partial_out / result_out
                    // (the pipeline's vector layout: counts at [4, 4 + 2K), {-, -, avg_acc, cnt, acc[K]}) when the NEXT step's
                    // last block - or hp_pipeline_flush_peer - collects them; the pending record lives in the workspace
                    peer_step_warp(link, ws, s_total, s_scratch, s_acc, K, defer ? partial_out : nullptr, defer ? result_out : s_result,
                                   defer, lane);
                    if (!defer) {
                        const bool bad = s_result[3] != s_result[3];  // poisoned by a timeout
                        for (int i = lane; i < 2 * K; i += 32) counts_out[i] = bad ? -1 : static_cast<int>(s_total[4 + i]);
                        for (int k = lane; k < K; k += 32) acc_out[k] = s_result[4 + k];
                        if (lane == 0) {
                            acc_out[K] = s_result[2];
                            acc_out[K + 1] = s_result[3];
                        }
                    }
                    if (lane == 0) ws->counter = 0;
                }
            } else {
                pck_publish(ws, K, counts_out, acc_out);
            }
        }
    }
}

// the staged-hi kernel covers: a high-resolution source, 32 column blocks (W == 128), 32 block rows (H == 128)
static bool block_hs_geometry(const FuseSrc& f, const BlockWalk& g, size_t& smem) {
    static const bool on = []() {
        const char* e = getenv("HP_FUSE_HS");
        return !(e && e[0] == '0');
    }();
    if (!on || f.hi == nullptr || g.cb != 32 || f.H != 4 * kBlockHsWarps * 4) return false;
    const size_t lo_b = 4ull * f.hl * f.wl, mid_b = f.mid ? 4ull * f.hm * f.wm : 0;
    const size_t rows_b = ((sizeof(RowTap) * 2 * static_cast<size_t>(f.H) + 127) / 128) * 128;
    smem = 2 * (lo_b + mid_b) + rows_b + static_cast<size_t>(kBlockHsWarps) * 2 * (2 * 4 * f.W * 4);
    return (2 * (lo_b + mid_b)) % 128 == 0 && smem <= 110 * 1024;  // 2 blocks per SM
}

// train1.py:410-424 in one launch: out [n, H, W] = a_lo up4(lo) + a_mid up2(mid), out2 [n, H/2, W/2] = a2 up2(lo);
// -> false when the geometry is not the 16 / 32 -> 64 (+ 32) pattern of the 4-warp block kernel with strips of two
static bool launch_fuse_pair(const FuseSrc& f, const BlockWalk& g, int sl, int sm, int n_warps, size_t smem, int n_maps, float* out,
                             float* out2, float a2, cudaStream_t s);

// block kernel applicable?  exact x2 / x4 scales, W/4 in {8, 16, 32}, bulk-copy alignment, everything fits
static bool block_geometry(const FuseSrc& f, const float* out, BlockWalk& g, int& sl, int& sm, int& n_warps, size_t& smem) {
    if (f.W % 4 != 0 || f.H % 4 != 0) return false;
    const int cb = f.W / 4;
    if (cb != 8 && cb != 16 && cb != 32) return false;
    if (f.hl < 2 || f.wl < 2) return false;
    if (f.hl * 4 == f.H && f.wl * 4 == f.W) sl = 4;
    else if (f.hl * 2 == f.H && f.wl * 2 == f.W) sl = 2;
    else return false;
    sm = 0;
    if (f.mid) {
        if (f.hm < 2 || f.wm < 2 || f.hm * 2 != f.H || f.wm * 2 != f.W) return false;
        sm = 2;
    }
    if (f.hi && !aligned16(f.hi)) return false;
    if (out && !aligned16(out)) return false;
    const size_t lo_b = 4ull * f.hl * f.wl, mid_b = f.mid ? 4ull * f.hm * f.wm : 0;
    if (lo_b % 16 != 0 || mid_b % 16 != 0 || !aligned16(f.lo) || (f.mid && !aligned16(f.mid))) return false;
    g.cb = cb;
    g.rps = 32 / cb;
    const int mb = f.H / 4;  // block rows
    n_warps = mb / g.rps;
    if (n_warps > kBlockMaxWarps) n_warps = kBlockMaxWarps;
    if (n_warps < 1 || mb % (n_warps * g.rps) != 0) return false;
    g.strip = mb / (n_warps * g.rps);
    smem = 2 * (lo_b + mid_b) + sizeof(RowTap) * 2 * static_cast<size_t>(f.H);
    return smem <= 48 * 1024;  // 4 blocks per SM
}

template <bool DECODE>
static bool launch_fuse_block(const FuseSrc& f, const BlockWalk& g, int sl, int sm, int n_warps, size_t smem, int n_maps, float* out,
                              const float* tgt_xy, int K, double thr, float* pred_xy, float* maxvals, int32_t* counts,
                              double* acc_out, Workspace* ws, cudaStream_t s, const PeerLink* link = nullptr, int defer = 0,
                              long long* partial_out = nullptr, double* result_out = nullptr) {
    // -> true when the launched kernel did the cross-GPU exchange of `link` itself (only the staged-hi kernel can)
    {
        size_t smem_hs = 0;
        if (sl == 4 && sm == 2 && block_hs_geometry(f, g, smem_hs)) {  // configs[3]: 32 / 64 / 128
            static bool configured[16] = {};
            int dev = 0;
            cudaGetDevice(&dev);
            if (dev < 0 || dev >= 16 || !configured[dev]) {
                cudaFuncSetAttribute(fuse_block_hs_kernel<4, 2, DECODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_hs));
                if (dev >= 0 && dev < 16) configured[dev] = true;
            }
            int sms = hp_device_sm_count();
            if (sms <= 0) sms = 148;
            const int grid_hs = n_maps < 2 * sms ? n_maps : 2 * sms;
            PeerLink lk{};
            lk.world = 1;
            if (link && DECODE) lk = *link;
            fuse_block_hs_kernel<4, 2, DECODE><<<grid_hs, 32 * kBlockHsWarps, smem_hs, s>>>(f, n_maps, out, tgt_xy, K, thr, pred_xy, maxvals,
                                                                                       counts, acc_out, ws, lk,
                                                                                       (link && DECODE) ? defer : 0, partial_out, result_out);
            return link != nullptr && DECODE;
        }
    }
    const int grid = rows_grid(n_maps);
    const int nt = 32 * n_warps;
    const char* shape = getenv("HP_FUSE_SHAPE");
    const bool dyn = shape && shape[0] == 'd';  // comparison runs: the run-time trip count (no unrolling by two)
#define HP_FUSE_BLOCK_S(SL, SM, STRIP)                                                                                            \
    fuse_block_kernel<SL, SM, DECODE, STRIP><<<grid, nt, smem, s>>>(f, g, n_maps, out, tgt_xy, K, thr, pred_xy, maxvals, counts, \
                                                                    acc_out, ws)
#define HP_FUSE_BLOCK(SL, SM)                          \
    do {                                               \
        if (g.strip == 2 && !dyn) HP_FUSE_BLOCK_S(SL, SM, 2); \
        else HP_FUSE_BLOCK_S(SL, SM, 0);               \
    } while (0)
    if (sl == 4 && sm == 2) HP_FUSE_BLOCK(4, 2);
    else if (sl == 4) HP_FUSE_BLOCK(4, 0);
    else if (sm == 2) HP_FUSE_BLOCK(2, 2);
    else HP_FUSE_BLOCK(2, 0);
#undef HP_FUSE_BLOCK_S
#undef HP_FUSE_BLOCK
    return false;
}

static bool launch_fuse_pair(const FuseSrc& f, const BlockWalk& g, int sl, int sm, int n_warps, size_t smem, int n_maps, float* out,
                             float* out2, float a2, cudaStream_t s) {
    if (sl != 4 || sm != 2 || g.strip != 2 || f.hi != nullptr || n_warps != kBlockMaxWarps) return false;
    if ((f.H / 8) * (f.W / 8) > 32 * n_warps || f.hl * 2 * 2 != f.H || f.wl * 2 * 2 != f.W || !aligned16(out2)) return false;
    fuse_block_kernel<4, 2, false, 2, true><<<rows_grid(n_maps), 32 * n_warps, smem, s>>>(f, g, n_maps, out, nullptr, 0, 0.0, nullptr, nullptr,
                                                                                     nullptr, nullptr, nullptr, out2, a2);
    return true;
}

}  // namespace hp
