// hp_pipeline_tiles.cuh - the production shape of the fused gen+loss+decode+PCK kernel.
//
// Work quantum = one TILE of 128*NV consecutive elements of one map (4 KB at NV = 8: a quarter of a
// 64x64 map, 1/16 of a 128x128 map, a whole 32x32 map).  Every WARP independently walks its tiles
// (static stride over the global tile index, so the 4 warps of a block stream the 4 quarters of the same
// map), double-buffered in registers so the next tile is always in flight, and publishes five numbers
// per tile - max, first index of the max, sum exp (relative to the max), sum p, sum p^2 - to a small
// global table, then bumps the map's arrival counter.  The warp that completes a map's LAST tile
// ("last arriver closes the door") merges the tiles, re-reads the <=169 patch pixels (L2), builds the
// target terms, decodes, scores PCK, closes MSE / KL and publishes.  No block barrier on the data path,
// no warp ever waits for another warp, any grid size balances to within one 4 KB tile.
//
// How it got here (profiles/r1_pipeline_*.md):
//   v1 block-per-map, target math in the hot loop ... 43 instructions/element, issue-bound (13 % of HBM)
//   v2 warp-per-map streaming ....................... 3x fewer instructions, but a 10 us work item on a
//                                                     2.3-wave grid leaves the SMs idle half the time
//   v3 4 warps per map, one barrier per map ......... every warp waits ~1200 cycles for the closing warp's
//                                                     scalar math at each barrier
//   all three ....................................... a single-thread last-block epilogue with ~50 dependent
//                                                     L2 round trips: a constant ~20 us tail
#pragma once
#include "hp_pipeline_common.cuh"

namespace hp {

constexpr int kTileWarps = 4;       // warps per block
constexpr int kTilesMaxPerMap = 4;  // more tiles per map than this -> warp-per-map streaming instead
constexpr int kTileMaxPatch = 6;    // patch pixels per lane of the closing warp: (2*tmp+1)^2 <= 192

struct TileStat {  // 32 bytes, one per (map, tile)
    float vmax;
    int idx;
    float s, sp, spp;
    float pad[3];
};

struct TileArgs {
    PipeArgs p;
    int tiles_per_map, n_tiles;
    FastDiv tdiv;             // by tiles_per_map
    TileStat* stats;          // [n_maps * tiles_per_map]   (workspace tail)
    unsigned int* arrivals;   // [n_maps], zero on entry, zero on exit
};

template <int NV>
__device__ __forceinline__ void tile_load(const float* __restrict__ pred, long long tile, int lane, float4 (&v)[NV]) {
    // maps are contiguous and HW is a multiple of the tile size, so tile t starts at t * 128 * NV
    const float4* p = reinterpret_cast<const float4*>(pred) + tile * (32 * NV) + lane;
#pragma unroll
    for (int j = 0; j < NV; ++j) v[j] = ldg_stream4(p + j * 32);
}

// hot loop over one tile -> TileStat (all lanes hold the result)
template <int NV, int LOSS>
__device__ __forceinline__ TileStat tile_stats(const float4 (&v)[NV], int tile_in_map, int lane) {
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
    TileStat r;
    r.vmax = warp_max(tm);
    int loc = 4 * NV;  // first index of the warp maximum inside this tile: scan downwards, lowest survives
#pragma unroll
    for (int j = NV - 1; j >= 0; --j) {
        loc = (v[j].w == r.vmax) ? (4 * j + 3) : loc;
        loc = (v[j].z == r.vmax) ? (4 * j + 2) : loc;
        loc = (v[j].y == r.vmax) ? (4 * j + 1) : loc;
        loc = (v[j].x == r.vmax) ? (4 * j + 0) : loc;
    }
    const int cand = (loc < 4 * NV) ? (tile_in_map * (128 * NV) + (loc >> 2) * 128 + lane * 4 + (loc & 3)) : 0x7fffffff;
    r.idx = warp_min_int(cand);
    r.s = 0.f;
    r.spp = 0.f;
    if (LOSS & HP_LOSS_KL) {
        const float ws = (r.vmax == -INFINITY) ? 0.0f : r.vmax;
        const float scale = (tm == -INFINITY) ? 0.0f : ex2_approx((tm - ws) * kLog2e);
        r.s = warp_sum((s2.x + s2.y) * scale);
    }
    r.sp = warp_sum(sp2.x + sp2.y);
    if (LOSS & HP_LOSS_MSE) r.spp = warp_sum(spp2.x + spp2.y);
    return r;
}

// per-warp exact loss accumulators (registers; every lane holds the same values)
struct WarpLoss {
    long long fx[2];
    int cls[6];
};
__device__ __forceinline__ void warp_loss_add(WarpLoss& w, int which, double v) {
    if (v != v) w.cls[3 * which + 0] += 1;
    else if (v >= kFxLimit) w.cls[3 * which + 1] += 1;
    else if (v <= -kFxLimit) w.cls[3 * which + 2] += 1;
    else w.fx[which] += __double2ll_rn(ldexp(v, kFxShift));
}

// the closing warp: merge the map's tiles, patch terms, decode, PCK, losses, publish
template <int LOSS>
__device__ __forceinline__ void tile_close_map(const TileArgs& t, int map, const TileStat& mine, int my_tile, int lane,
                                               const float* s_tab, WarpLoss& wl) {
    const PipeArgs& a = t.p;
    const float* pm = a.pred + static_cast<size_t>(map) * a.HW;
    float weight;
    const Centre c = pipe_centre(a, a.joints[2 * map], a.joints[2 * map + 1], a.vis[map], weight);
    const bool pasted = c.y != kNoPaste;
    // patch pixels: issue the loads first, merge the tiles while they fly
    float pv[kTileMaxPatch], tv[kTileMaxPatch];
    {
        const int side = 2 * a.tmp + 1, n_patch = side * side;
#pragma unroll
        for (int k = 0; k < kTileMaxPatch; ++k) {
            const int i = lane + 32 * k;
            uint32_t ry, rx;
            a.sdiv.divmod(static_cast<uint32_t>(i), ry, rx);
            const int dx = static_cast<int>(rx) - a.tmp, dy = static_cast<int>(ry) - a.tmp;
            const int x = c.x + dx, y = c.y + dy;
            const bool in = pasted && i < n_patch && x >= 0 && x < a.W && y >= 0 && y < a.H;
            pv[k] = in ? ldg_stream1(pm + y * a.W + x) : 0.0f;
            tv[k] = in ? s_tab[dx * dx + dy * dy] : 0.0f;
        }
    }
    // merge: lane l takes tiles l, l+32, ... (other warps' rows come from L2, mine from registers)
    ArgMax am{-INFINITY, 0x7fffffff};
    float m_run = -INFINITY, s_run = 0.f, sp = 0.f, spp = 0.f;
    const volatile TileStat* row = t.stats + static_cast<size_t>(map) * t.tiles_per_map;
    for (int q = lane; q < t.tiles_per_map; q += 32) {
        TileStat ts;
        if (q == my_tile) {
            ts = mine;
        } else {
            ts.vmax = row[q].vmax;
            ts.idx = row[q].idx;
            ts.s = row[q].s;
            ts.sp = row[q].sp;
            ts.spp = row[q].spp;
        }
        if (ts.vmax > am.v || (ts.vmax == am.v && ts.idx < am.i)) {
            am.v = ts.vmax;
            am.i = ts.idx;
        }
        if (LOSS & HP_LOSS_KL) {
            const float mn = fmaxf(m_run, ts.vmax);
            const float ms = (mn == -INFINITY) ? 0.0f : mn;
            s_run = s_run * ((m_run == -INFINITY) ? 0.0f : ex2_approx((m_run - ms) * kLog2e)) +
                    ts.s * ((ts.vmax == -INFINITY) ? 0.0f : ex2_approx((ts.vmax - ms) * kLog2e));
            m_run = mn;
        }
        sp += ts.sp;
        spp += ts.spp;
    }
    am = warp_argmax(am, lane);  // equal maxima: the lower index (earlier tile) wins
    float sum_exp = 0.f;
    if (LOSS & HP_LOSS_KL) {
        const float Ms = (am.v == -INFINITY) ? 0.0f : am.v;
        sum_exp = warp_sum(s_run * ((m_run == -INFINITY) ? 0.0f : ex2_approx((m_run - Ms) * kLog2e)));
    }
    const float sum_p = warp_sum(sp);
    const float sum_pp = (LOSS & HP_LOSS_MSE) ? warp_sum(spp) : 0.f;
    if (sum_p != sum_p) {
        // a NaN (or +inf with -inf) is in the map: redo the argmax with numpy's exact rules from memory
        ArgMax sx = am_init();
        const float4* m4 = reinterpret_cast<const float4*>(pm);
        for (int e4 = lane; e4 < a.HW / 4; e4 += 32) am_scan4<true>(sx, ldg_stream4(m4 + e4), e4 * 4);
        am = warp_argmax(sx, lane);
        sum_exp = __int_as_float(0x7fc00000);  // log_softmax of a map holding a NaN is NaN
    }
    PatchSums ps{0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < kTileMaxPatch; ++k) patch_pixel<LOSS>(ps, tv[k], pv[k], a.eps);
    if (LOSS & HP_LOSS_KL) {
        ps.up = warp_sum(ps.up);
        ps.ulogu = warp_sum(ps.ulogu);
        ps.u = warp_sum(ps.u);
        ps.p = warp_sum(ps.p);
    }
    if (LOSS & HP_LOSS_MSE) ps.e = warp_sum(ps.e);

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
    if (LOSS & HP_LOSS_MSE) warp_loss_add(wl, 0, mse);
    if (LOSS & HP_LOSS_KL) warp_loss_add(wl, 1, kl);
    if (lane == 0) {
        a.pred_xy[2 * map + 0] = px;
        a.pred_xy[2 * map + 1] = py;
        if (a.maxvals) a.maxvals[map] = am.v;
        if (a.weight_out) a.weight_out[map] = weight;
        const int k = map % a.K;
        if (valid) atomicAdd(&a.ws->counts[a.K + k], 1);
        if (hit) atomicAdd(&a.ws->counts[k], 1);
        if (t.tiles_per_map > 1) t.arrivals[map] = 0;  // leave the workspace zeroed for the next launch
    }
}

// publish a tile's statistics and learn whether this warp closed the map (all lanes get the answer)
__device__ __forceinline__ unsigned int tile_arrive(const TileArgs& t, int map, int tile_in_map, const TileStat& st,
                                                    int lane) {
    unsigned int prev = 0;
    if (lane == 0) {
        TileStat* dst = t.stats + static_cast<size_t>(map) * t.tiles_per_map + tile_in_map;
        *reinterpret_cast<float4*>(dst) = make_float4(st.vmax, __int_as_float(st.idx), st.s, st.sp);
        dst->spp = st.spp;
        __threadfence();  // statistics visible before the arrival is
        prev = atomicAdd(&t.arrivals[map], 1u);
    }
    return __shfl_sync(0xffffffffu, prev, 0);
}

template <int NV, int LOSS>
__global__ void __launch_bounds__(32 * kTileWarps, 4) pipeline_tiles_kernel(const TileArgs t) {
    extern __shared__ float s_tab[];
    const PipeArgs& a = t.p;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long n_warps = static_cast<long long>(gridDim.x) * kTileWarps;
    long long tile = static_cast<long long>(blockIdx.x) * kTileWarps + warp;

    float4 bufA[NV], bufB[NV];
    if (tile < t.n_tiles) tile_load<NV>(a.pred, tile, lane, bufA);
    load_table(s_tab, a.tab, a.tmp);
    __syncthreads();  // the only block barrier before the epilogue

    WarpLoss wl;
    wl.fx[0] = wl.fx[1] = 0;
#pragma unroll
    for (int i = 0; i < 6; ++i) wl.cls[i] = 0;
    const FastDiv tdiv = t.tdiv;

    // two tiles per trip so the register buffers have static names
    while (tile < t.n_tiles) {
        long long next = tile + n_warps;
        if (next < t.n_tiles) tile_load<NV>(a.pred, next, lane, bufB);
        {
            uint32_t map, q;
            tdiv.divmod(static_cast<uint32_t>(tile), map, q);
            const TileStat st = tile_stats<NV, LOSS>(bufA, static_cast<int>(q), lane);
            const bool closes = (t.tiles_per_map == 1) ||
                                (tile_arrive(t, static_cast<int>(map), static_cast<int>(q), st, lane) ==
                                 static_cast<unsigned int>(t.tiles_per_map - 1));
            if (closes) {
                if (t.tiles_per_map > 1) __threadfence();  // acquire: the other tiles' statistics
                tile_close_map<LOSS>(t, static_cast<int>(map), st, static_cast<int>(q), lane, s_tab, wl);
            }
        }
        tile = next;
        if (tile >= t.n_tiles) break;
        next = tile + n_warps;
        if (next < t.n_tiles) tile_load<NV>(a.pred, next, lane, bufA);
        {
            uint32_t map, q;
            tdiv.divmod(static_cast<uint32_t>(tile), map, q);
            const TileStat st = tile_stats<NV, LOSS>(bufB, static_cast<int>(q), lane);
            const bool closes = (t.tiles_per_map == 1) ||
                                (tile_arrive(t, static_cast<int>(map), static_cast<int>(q), st, lane) ==
                                 static_cast<unsigned int>(t.tiles_per_map - 1));
            if (closes) {
                if (t.tiles_per_map > 1) __threadfence();
                tile_close_map<LOSS>(t, static_cast<int>(map), st, static_cast<int>(q), lane, s_tab, wl);
            }
        }
        tile = next;
    }
    // ---- epilogue: exact loss sums -> workspace, last block publishes -------------------------------------
    if (lane == 0) {
        for (int w = 0; w < 2; ++w)
            if (wl.fx[w] != 0) atomicAdd(&a.ws->acc[w], static_cast<unsigned long long>(wl.fx[w]));
        for (int i = 0; i < 6; ++i)
            if (wl.cls[i] != 0) atomicAdd(&a.ws->acc[2 + i], static_cast<unsigned long long>(wl.cls[i]));
        __threadfence();
    }
    if (pipeline_last_block(a.ws)) pipeline_publish(a);
}

}  // namespace hp
