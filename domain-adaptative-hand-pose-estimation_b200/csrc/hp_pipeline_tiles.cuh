// hp_pipeline_tiles.cuh - the production shape of the fused gen+loss+decode+PCK kernel.
//
// Work quantum = one TILE of 128*NV consecutive elements of one map (4 KB at NV = 8: a quarter of a
// 64x64 map, a whole 32x32 map).  Every WARP independently walks its tiles with a static stride over
// the global tile index - which makes the 4 warps of a block stream the 4 quarters of the SAME map -
// double-buffered in registers so the next tile is always in flight, and leaves four numbers per tile
// (max, sum exp relative to the max, sum p, sum p^2) in a shared-memory ring slot, then bumps the
// slot's arrival counter (shared-memory RED, block-scope fence: no global fence or atomic on the data
// path).  Closing duty ROTATES over the warps: map j of a block is closed by warp j mod 4 as soon as it
// sees all arrivals (polled after each of its own tiles).  Closing = merge the tiles, re-read from L2
// the one tile that holds the maximum (first-index scan) and the <=169 patch pixels, build the target
// terms, decode, score PCK, close MSE / KL, publish.  No block barrier in the loop, no warp waits for
// another (the ring is 8 maps deep), and the grid balances to within one 4 KB tile.
//
// How it got here (profiles/r1_pipeline_history.md):
//   v1 block-per-map, target math in the hot loop ... 43 instructions/element, issue-bound (13 % of HBM)
//   v2 warp-per-map streaming ....................... 3x fewer instructions, but a 10 us work item on a
//                                                     2.3-wave grid leaves the SMs idle half the time
//   v3 4 warps per map, one barrier per map ......... every warp waits ~1200 cycles for the closing warp's
//                                                     scalar math at each barrier
//   v1-v3 ........................................... a single-thread last-block epilogue with ~50 dependent
//                                                     L2 round trips: a constant ~20 us tail
//   v4 last arriver through GLOBAL counters ......... a __threadfence + atomic round trip per tile with 31
//                                                     lanes parked at the reconvergence point (20 % issue)
//   v5 last arriver closes, shared-memory ring ...... the slowest warp is always last, closes every map and
//                                                     falls further behind; the other three idle at the end
//   v6 rotating closers ............................. 2730 instructions/map: spills at 127 registers, an eager
//                                                     index scan per tile, per-pixel patch index arithmetic
//   v7 (this) ....................................... index scan only by the closer, per-lane patch table in
//                                                     shared memory, transposed multi-value reductions
#pragma once
#include "hp_pipeline_common.cuh"

namespace hp {

constexpr int kTileWarps = 4;       // warps per block
constexpr int kTilesMaxPerMap = 4;  // tiles_per_map in {1, 2, 4}; larger maps use the stream shape
constexpr int kTileMaxPatch = 6;    // patch pixels per lane of the closing warp: (2*tmp+1)^2 <= 192
constexpr int kTileRing = 8;        // ring depth (maps in flight per block) of the shared-memory slots

struct TileStat {  // what one warp leaves per tile
    float vmax, s, sp, spp;
};

struct TileArgs {
    PipeArgs p;
    int tiles_per_map, n_tiles;  // the warps of a block split into 4/tpm groups, one map per group
    FastDiv tdiv;                // by tiles_per_map
    FastDiv kdiv;                // by K (joint index of a map)
};

struct TileRing {
    TileStat stat[kTileRing][kTileWarps];
    unsigned int count[kTileRing][kTileWarps];  // arrivals per (slot, group)
    unsigned int gen[kTileRing][kTileWarps];    // how many times (slot, group) has been closed
};

// per-lane patch slots: offset from the centre and the target terms that do not depend on the prediction
struct PatchSlot {
    int dx, dy;        // dx = 1<<20 for unused slots (never in bounds)
    float t, ulogu;    // target value, (t+eps)*ln(t+eps)
};

template <int NV>
__device__ __forceinline__ void tile_load(const float* __restrict__ pred, int tile, int lane, float4 (&v)[NV]) {
    // maps are contiguous and HW is a multiple of the tile size, so tile t starts at t * 128 * NV
    const float4* p = reinterpret_cast<const float4*>(pred) + static_cast<size_t>(tile) * (32 * NV) + lane;
#pragma unroll
    for (int j = 0; j < NV; ++j) v[j] = ldg_stream4(p + j * 32);
}

// Sum three values over the warp with 6 shuffles instead of 15: after each exchange a lane keeps half of
// its values.  On return lanes 0-7 hold sum(a), lanes 8-15 sum(b), lanes 16-23 sum(c), lanes 24-31 zero.
__device__ __forceinline__ float warp_sum3_scattered(float a, float b, float c, int lane) {
    const bool hi16 = (lane & 16) != 0;
    float k0 = hi16 ? c : a, k1 = hi16 ? 0.0f : b;  // low half keeps (a, b), high half keeps (c, 0)
    const float s0 = hi16 ? a : c, s1 = hi16 ? b : 0.0f;
    k0 += __shfl_xor_sync(0xffffffffu, s0, 16);
    k1 += __shfl_xor_sync(0xffffffffu, s1, 16);
    const bool hi8 = (lane & 8) != 0;
    float k = hi8 ? k1 : k0;
    const float s = hi8 ? k0 : k1;
    k += __shfl_xor_sync(0xffffffffu, s, 8);
    k += __shfl_xor_sync(0xffffffffu, k, 4);
    k += __shfl_xor_sync(0xffffffffu, k, 2);
    k += __shfl_xor_sync(0xffffffffu, k, 1);
    return k;
}

// hot loop over one tile; vmax in every lane, the three sums scattered (see warp_sum3_scattered)
template <int NV, int LOSS>
__device__ __forceinline__ void tile_stats(const float4 (&v)[NV], int lane, float& vmax, float& scattered) {
    float tm = -INFINITY;
#pragma unroll
    for (int j = 0; j < NV; ++j) tm = fmaxf(tm, fmaxf(fmaxf(v[j].x, v[j].y), fmaxf(v[j].z, v[j].w)));
    float2 s2 = make_float2(0.f, 0.f), sp2 = make_float2(0.f, 0.f), spp2 = make_float2(0.f, 0.f);
    if (LOSS & HP_LOSS_KL) {
        const float ms = (tm == -INFINITY) ? 0.0f : tm;
        const float2 l2 = make_float2(kLog2e, kLog2e), mb2 = make_float2(-ms * kLog2e, -ms * kLog2e);
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            const float2 a0 = __ffma2_rn(make_float2(v[j].x, v[j].y), l2, mb2);
            const float2 a1 = __ffma2_rn(make_float2(v[j].z, v[j].w), l2, mb2);
            s2 = __fadd2_rn(s2, make_float2(ex2_approx(a0.x), ex2_approx(a0.y)));
            s2 = __fadd2_rn(s2, make_float2(ex2_approx(a1.x), ex2_approx(a1.y)));
        }
    }
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        const float2 lo = make_float2(v[j].x, v[j].y), hi = make_float2(v[j].z, v[j].w);
        sp2 = __fadd2_rn(sp2, __fadd2_rn(lo, hi));
        if (LOSS & HP_LOSS_MSE) {
            spp2 = __ffma2_rn(lo, lo, spp2);
            spp2 = __ffma2_rn(hi, hi, spp2);
        }
    }
    vmax = warp_max(tm);
    float s = 0.f;
    if (LOSS & HP_LOSS_KL) {
        const float ws = (vmax == -INFINITY) ? 0.0f : vmax;
        s = (s2.x + s2.y) * ((tm == -INFINITY) ? 0.0f : ex2_approx((tm - ws) * kLog2e));
    }
    scattered = warp_sum3_scattered(s, sp2.x + sp2.y, spp2.x + spp2.y, lane);
}

// per-warp exact loss accumulators (shared memory, touched by lane 0 of the owning warp only)
struct WarpLoss {
    long long fx[2];
    int cls[6];
};
static __device__ __noinline__ void warp_loss_add_nonfinite(WarpLoss* w, int which, double v) {
    if (v != v) w->cls[3 * which + 0] += 1;
    else if (v >= kFxLimit) w->cls[3 * which + 1] += 1;
    else w->cls[3 * which + 2] += 1;
}
__device__ __forceinline__ void warp_loss_add(WarpLoss* w, int which, double v) {
    if (fabs(v) < kFxLimit) w->fx[which] += __double2ll_rn(v * 1099511627776.0);  // * 2^40, exact scaling
    else warp_loss_add_nonfinite(w, which, v);
}

// the closing warp: merge the map's tiles, index scan, patch terms, decode, PCK, losses, publish
template <int NV, int LOSS>
__device__ __forceinline__ void tile_close_map(const TileArgs& t, int map, TileRing* ring, int slot, int g,
                                               unsigned int next_gen, int lane, const PatchSlot* s_patch,
                                               WarpLoss* wl) {
    const PipeArgs& a = t.p;
    const int tpm = t.tiles_per_map;
    const float* pm = a.pred + static_cast<size_t>(map) * a.HW;
    const volatile TileStat* st = &ring->stat[slot][g * tpm];
    // which tile holds the maximum?  a strict > keeps the earlier tile - and so the lower indices - on ties
    float M = st[0].vmax;
    int qstar = 0;
#pragma unroll
    for (int q = 1; q < kTilesMaxPerMap; ++q)
        if (q < tpm) {
            const float vq = st[q].vmax;
            if (vq > M) {
                M = vq;
                qstar = q;
            }
        }
    // everything that has to come from memory is requested now, in one go: the tile holding the maximum
    // (L2, it was streamed microseconds ago), the patch pixels, the keypoint
    float4 w[NV];
    tile_load<NV>(pm, qstar, lane, w);
    float weight;
    const Centre c = pipe_centre(a, a.joints[2 * map], a.joints[2 * map + 1], a.vis[map], weight);
    const bool pasted = c.y != kNoPaste;
    // patch terms with u = t + eps:  up = sum u p, u = sum u, p = sum p, e = sum t (t - 2p), ulogu = sum u ln u
    PatchSums ps{0.f, 0.f, 0.f, 0.f, 0.f};
    float pv[kTileMaxPatch], tv[kTileMaxPatch];
#pragma unroll
    for (int k = 0; k < kTileMaxPatch; ++k) {
        const PatchSlot sl = s_patch[k * 32 + lane];
        const int x = c.x + sl.dx, y = c.y + sl.dy;
        const bool in = pasted && static_cast<unsigned>(x) < static_cast<unsigned>(a.W) &&
                        static_cast<unsigned>(y) < static_cast<unsigned>(a.H);
        pv[k] = in ? ldg_stream1(pm + y * a.W + x) : 0.0f;
        tv[k] = in ? sl.t : 0.0f;
        if (LOSS & HP_LOSS_KL) {  // the part that does not need the prediction
            ps.ulogu += in ? sl.ulogu : 0.0f;
            ps.u += (tv[k] != 0.0f) ? tv[k] + a.eps : 0.0f;
        }
    }
    // merge the per-tile sums while the loads fly, then hand the ring slot back
    float sum_exp = 0.f, sum_p = 0.f, sum_pp = 0.f;
    const float Ms = (M == -INFINITY) ? 0.0f : M;
#pragma unroll
    for (int q = 0; q < kTilesMaxPerMap; ++q)
        if (q < tpm) {
            if (LOSS & HP_LOSS_KL) {
                const float vq = st[q].vmax;
                sum_exp += st[q].s * ((vq == -INFINITY) ? 0.0f : ex2_approx((vq - Ms) * kLog2e));
            }
            sum_p += st[q].sp;
            if (LOSS & HP_LOSS_MSE) sum_pp += st[q].spp;
        }
    __syncwarp();
    if (lane == 0) {
        ring->count[slot][g] = 0;
        __threadfence_block();
        ring->gen[slot][g] = next_gen;
    }
    // first index of M inside tile qstar: scan downwards so the lowest survives, then min over the lanes
    ArgMax am;
    am.v = M;
    {
        int loc = 4 * NV;
#pragma unroll
        for (int j = NV - 1; j >= 0; --j) {
            loc = (w[j].w == M) ? (4 * j + 3) : loc;
            loc = (w[j].z == M) ? (4 * j + 2) : loc;
            loc = (w[j].y == M) ? (4 * j + 1) : loc;
            loc = (w[j].x == M) ? (4 * j + 0) : loc;
        }
        const int cand = (loc < 4 * NV) ? (qstar * (128 * NV) + (loc >> 2) * 128 + lane * 4 + (loc & 3)) : 0x7fffffff;
        am.i = warp_min_int(cand);
    }
    if (sum_p != sum_p) {
        // a NaN (or +inf with -inf) is in the map: redo the argmax with numpy's exact rules from memory
        ArgMax sx = am_init();
        const float4* m4 = reinterpret_cast<const float4*>(pm);
        for (int e4 = lane; e4 < a.HW / 4; e4 += 32) am_scan4<true>(sx, ldg_stream4(m4 + e4), e4 * 4);
        am = warp_argmax(sx, lane);
        sum_exp = __int_as_float(0x7fc00000);  // log_softmax of a map holding a NaN is NaN
    }
#pragma unroll
    for (int k = 0; k < kTileMaxPatch; ++k) {
        const float tk = tv[k], pk = pv[k];
        if (LOSS & HP_LOSS_KL) {
            const float u = (tk != 0.0f) ? tk + a.eps : 0.0f;
            ps.up = fmaf(u, pk, ps.up);
            ps.p += pk;
        }
        if (LOSS & HP_LOSS_MSE) ps.e = fmaf(tk, tk - 2.0f * pk, ps.e);
    }
    if (LOSS & HP_LOSS_KL) {
        const float r = warp_sum3_scattered(ps.up, ps.u, ps.p, lane);
        ps.up = __shfl_sync(0xffffffffu, r, 0);
        ps.u = __shfl_sync(0xffffffffu, r, 8);
        ps.p = __shfl_sync(0xffffffffu, r, 16);
    }
    {
        const float r = warp_sum3_scattered(ps.ulogu, ps.e, 0.0f, lane);
        ps.ulogu = __shfl_sync(0xffffffffu, r, 0);
        ps.e = __shfl_sync(0xffffffffu, r, 8);
    }

    uint32_t qy, qx;
    a.wdiv.divmod(static_cast<uint32_t>(am.i), qy, qx);
    const float keep = (am.v > 0.0f) ? 1.0f : 0.0f;  // NaN -> 0 (keypoint_detection.py:31-34)
    const float px = static_cast<float>(qx) * keep, py = static_cast<float>(qy) * keep;
    // decoding the generated target: its unique maximum (exactly 1.0) sits on the centre when pasted;
    // the all-zero map decodes to the masked (0,0)   
    const float tx = pasted ? static_cast<float>(c.x) : 0.0f, ty = pasted ? static_cast<float>(c.y) : 0.0f;
    int valid, hit;
    pipe_pck(a, px, py, tx, ty, valid, hit);
    double mse, kl;
    pipe_losses<LOSS>(a, c, weight, am.v, sum_exp, sum_p, sum_pp, ps, mse, kl);
    if (lane == 0) {
        if (LOSS & HP_LOSS_MSE) warp_loss_add(wl, 0, mse);
        if (LOSS & HP_LOSS_KL) warp_loss_add(wl, 1, kl);
        a.pred_xy[2 * map + 0] = px;
        a.pred_xy[2 * map + 1] = py;
        if (a.maxvals) a.maxvals[map] = am.v;
        if (a.weight_out) a.weight_out[map] = weight;
        const int k = map - static_cast<int>(t.kdiv.div(static_cast<uint32_t>(map))) * a.K;
        if (valid) atomicAdd(&a.ws->counts[a.K + k], 1);
        if (hit) atomicAdd(&a.ws->counts[k], 1);
    }
}

// Close every map this warp is the designated closer of and whose tiles have all arrived.
// Designation rotates: within its group of tpm warps, warp r closes the iterations j with j % tpm == r,
// so closing work is spread evenly no matter which warp happens to arrive last (a last-arriver rule
// makes the slowest warp close every map and fall ever further behind).
template <int NV, int LOSS>
__device__ __forceinline__ void tile_try_close(const TileArgs& t, int& next_close, int upto, bool drain,
                                               int tile0, int n_warps, int tpm, int g, int lane,
                                               const PatchSlot* s_patch, TileRing* ring, WarpLoss* wl) {
    while (next_close <= upto) {
        const int slot = next_close & (kTileRing - 1);
        const volatile unsigned int* cnt = &ring->count[slot][g];
        if (*cnt != static_cast<unsigned int>(tpm)) {
            if (!drain) return;
            __nanosleep(1000);  // end of the walk: the other warps of the group are still on their way
            continue;
        }
        __threadfence_block();
        // the group's tiles in iteration j start at tile0_of_block + j*n_warps + g*tpm, all of one map
        const int first_tile = tile0 + next_close * n_warps + g * tpm;
        const int map = static_cast<int>(t.tdiv.div(static_cast<uint32_t>(first_tile)));
        tile_close_map<NV, LOSS>(t, map, ring, slot, g, static_cast<unsigned int>(next_close / kTileRing) + 1u, lane,
                                 s_patch, wl);
        next_close += tpm;
    }
}

// one tile: statistics -> ring slot -> arrival (fire and forget)
template <int NV, int LOSS>
__device__ __forceinline__ void tile_step(const float4 (&v)[NV], int iter, int warp, int g, int lane, TileRing* ring) {
    float vmax, scattered;
    tile_stats<NV, LOSS>(v, lane, vmax, scattered);
    // the warps [g*tpm, (g+1)*tpm) of this block hold the tiles of the same map in this iteration
    const int slot = iter & (kTileRing - 1);
    if (iter >= kTileRing) {  // the slot was used 8 maps ago: make sure that map has been read out
        const unsigned int my_gen = static_cast<unsigned int>(iter) / kTileRing;
        const volatile unsigned int* gen = &ring->gen[slot][g];
        while (*gen != my_gen) {  // practically never taken
        }
    }
    // lanes 0 / 8 / 16 hold sum exp / sum p / sum p^2 (warp_sum3_scattered); lane 24 carries the max
    float* dst = &ring->stat[slot][warp].vmax;
    if ((lane & 7) == 0) dst[lane == 24 ? 0 : 1 + (lane >> 3)] = (lane == 24) ? vmax : scattered;
    __syncwarp();
    if (lane == 0) {
        __threadfence_block();  // statistics visible (block scope) before the arrival is
        atomicAdd(&ring->count[slot][g], 1u);
    }
}

// shared prologue: ring + per-lane patch table (slot k of lane l is patch pixel i = l + 32k, row-major)
__device__ __forceinline__ void tile_prologue(const PipeArgs& a, TileRing* ring, PatchSlot* s_patch) {
    for (int i = threadIdx.x; i < kTileRing * kTileWarps; i += blockDim.x) {
        (&ring->count[0][0])[i] = 0;
        (&ring->gen[0][0])[i] = 0;
    }
    const int side = 2 * a.tmp + 1, n_patch = side * side;
    for (int i = threadIdx.x; i < kTileMaxPatch * 32; i += blockDim.x) {
        PatchSlot ps;
        ps.dx = 1 << 20;
        ps.dy = 0;
        ps.t = 0.f;
        ps.ulogu = 0.f;
        if (i < n_patch) {
            uint32_t ry, rx;
            a.sdiv.divmod(static_cast<uint32_t>(i), ry, rx);
            ps.dx = static_cast<int>(rx) - a.tmp;
            ps.dy = static_cast<int>(ry) - a.tmp;
            ps.t = a.tab[ps.dx * ps.dx + ps.dy * ps.dy];
            const float u = ps.t + a.eps;
            ps.ulogu = (u != 0.0f) ? u * logf(u) : 0.0f;
        }
        s_patch[i] = ps;
    }
}

// shared epilogue: exact loss sums -> workspace, last block publishes
__device__ __forceinline__ void tile_epilogue(const PipeArgs& a, const WarpLoss* wl, int lane) {
    if (lane == 0) {
        for (int w = 0; w < 2; ++w)
            if (wl->fx[w] != 0) atomicAdd(&a.ws->acc[w], static_cast<unsigned long long>(wl->fx[w]));
        for (int i = 0; i < 6; ++i)
            if (wl->cls[i] != 0) atomicAdd(&a.ws->acc[2 + i], static_cast<unsigned long long>(wl->cls[i]));
        __threadfence();
    }
    if (pipeline_last_block(a.ws)) pipeline_publish(a);
}

// Variant A: two register buffers (the next tile is requested before the current one is reduced),
// 4 blocks = 16 warps per SM.
template <int NV, int LOSS>
__global__ void __launch_bounds__(32 * kTileWarps, 4) pipeline_tiles_kernel(const TileArgs t) {
    __shared__ TileRing s_ring;
    __shared__ PatchSlot s_patch[kTileMaxPatch * 32];
    __shared__ WarpLoss s_wl[kTileWarps];
    const PipeArgs& a = t.p;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_warps = static_cast<int>(gridDim.x) * kTileWarps;
    const int tile0 = static_cast<int>(blockIdx.x) * kTileWarps;
    int tile = tile0 + warp;

    float4 bufA[NV], bufB[NV];
    if (tile < t.n_tiles) tile_load<NV>(a.pred, tile, lane, bufA);
    tile_prologue(a, &s_ring, s_patch);
    if (lane == 0) {
        s_wl[warp].fx[0] = s_wl[warp].fx[1] = 0;
        for (int i = 0; i < 6; ++i) s_wl[warp].cls[i] = 0;
    }
    __syncthreads();  // the only block barrier before the epilogue

    WarpLoss* wl = &s_wl[warp];
    const int tpm = t.tiles_per_map;              // 1, 2 or 4
    const int g = tpm == 4 ? 0 : (tpm == 2 ? (warp >> 1) : warp);
    int next_close = warp & (tpm - 1);  // first iteration this warp is the designated closer of

    // Two tiles per trip so the register buffers have static names; ONE closing call site per trip so the
    // long scalar closure is inlined once (it would otherwise triple the code and spill the tile buffers).
    int iter = 0;
    bool more = tile < t.n_tiles;
    while (true) {
        if (more) {
            int next = tile + n_warps;
            if (next < t.n_tiles) tile_load<NV>(a.pred, next, lane, bufB);
            tile_step<NV, LOSS>(bufA, iter, warp, g, lane, &s_ring);
            ++iter;
            tile = next;
            more = tile < t.n_tiles;
            if (more) {
                next = tile + n_warps;
                if (next < t.n_tiles) tile_load<NV>(a.pred, next, lane, bufA);
                tile_step<NV, LOSS>(bufB, iter, warp, g, lane, &s_ring);
                ++iter;
                tile = next;
                more = tile < t.n_tiles;
            }
        }
        // close what is ready; after the last tile, wait for the rest of this warp's maps (drain)
        tile_try_close<NV, LOSS>(t, next_close, iter - 1, !more, tile0, n_warps, tpm, g, lane, s_patch, &s_ring, wl);
        if (!more) break;
    }
    tile_epilogue(a, wl, lane);
}

// Variant B: one register buffer, memory latency hidden by residency instead of register prefetch:
// <= 80 registers, 6 blocks = 24 warps per SM.
template <int NV, int LOSS>
__device__ __forceinline__ void pipeline_tiles1_body(const TileArgs& t) {
    __shared__ TileRing s_ring;
    __shared__ PatchSlot s_patch[kTileMaxPatch * 32];
    __shared__ WarpLoss s_wl[kTileWarps];
    const PipeArgs& a = t.p;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_warps = static_cast<int>(gridDim.x) * kTileWarps;
    const int tile0 = static_cast<int>(blockIdx.x) * kTileWarps;
    int tile = tile0 + warp;

    float4 buf[NV];
    if (tile < t.n_tiles) tile_load<NV>(a.pred, tile, lane, buf);
    tile_prologue(a, &s_ring, s_patch);
    if (lane == 0) {
        s_wl[warp].fx[0] = s_wl[warp].fx[1] = 0;
        for (int i = 0; i < 6; ++i) s_wl[warp].cls[i] = 0;
    }
    __syncthreads();

    WarpLoss* wl = &s_wl[warp];
    const int tpm = t.tiles_per_map;
    const int g = tpm == 4 ? 0 : (tpm == 2 ? (warp >> 1) : warp);
    int next_close = warp & (tpm - 1);
    int iter = 0;
    bool more = tile < t.n_tiles;
    while (true) {
        if (more) {
            if (iter > 0) tile_load<NV>(a.pred, tile, lane, buf);
            tile_step<NV, LOSS>(buf, iter, warp, g, lane, &s_ring);
            ++iter;
            tile += n_warps;
            more = tile < t.n_tiles;
        }
        tile_try_close<NV, LOSS>(t, next_close, iter - 1, !more, tile0, n_warps, tpm, g, lane, s_patch, &s_ring, wl);
        if (!more) break;
    }
    tile_epilogue(a, wl, lane);
}
template <int NV, int LOSS>
__global__ void __launch_bounds__(32 * kTileWarps, 6) pipeline_tiles1_kernel(const TileArgs t) {
    pipeline_tiles1_body<NV, LOSS>(t);
}
// Variant C: the same single-buffer body at 5 blocks = 20 warps per SM (<= 102 registers)
template <int NV, int LOSS>
__global__ void __launch_bounds__(32 * kTileWarps, 5) pipeline_tiles1c_kernel(const TileArgs t) {
    pipeline_tiles1_body<NV, LOSS>(t);
}

}  // namespace hp
