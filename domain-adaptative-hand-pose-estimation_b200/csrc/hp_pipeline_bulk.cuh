// hp_pipeline_bulk.cuh - the production shape of the fused gen+loss+decode+PCK kernel on sm_100a:
// maps are staged in shared memory by the TMA engine (cp.async.bulk + mbarrier complete_tx), one
// persistent block per SM, every warp owns a PRIVATE ring of stages.
//
// Why this shape (profiles/r1_pipeline_history.md): the register-tile kernels needed the warps themselves
// to keep HBM busy, so memory-level parallelism was tied to occupancy and to 128-register tile buffers
// (spills), and a map's four tiles had to be merged through a cross-warp ring.  Here
//   * bytes in flight are decoupled from the instruction stream: the copy engine keeps W*KST stages
//     (12 x 16 KB = 192 KB per SM) requested, the warps only ever touch shared memory;
//   * a warp owns a whole map: no cross-warp protocol, no block barrier, no atomics in the loop; the
//     stage's barrier is armed and re-filled by the same warp that drained it (no empty-barrier, no
//     phase hazards between warps);
//   * the map sits in shared memory, so a two-pass exact softmax (max+argmax, then sums against the true
//     max), the first-index resolution and the <=169 patch pixels are cheap LDS traffic, not L2 re-reads;
//   * everything that does not need the prediction (centre, patch offsets, target-only sums) runs before
//     the wait on the stage, the float64 closure after the refill has been issued.
// Large maps (128x128 = 64 KB) go through the same loop in 16 KB chunks with an online softmax merge.
//
// Algorithmic bytes per map: H*W*4 + 24 + 8; the map is read from HBM exactly once.
#pragma once
#include "hp_internal.cuh"
#include "hp_pipeline_common.cuh"
#include "hp_peer_step.cuh"
#include "hp_tma.cuh"
#include "hp_pipeline_parts.cuh"  // PatchSlot, WarpLoss, warp_sum3_scattered, kTileMaxPatch

namespace hp {

struct BulkArgs {
    PipeArgs p;
    FastDiv kdiv;   // by K (joint index of a map)
    int n_chunks;   // chunks per map (1 unless the map is larger than a stage)
    int overlap;    // 0: serialised launch; d >= 1: programmatic dependent launch on 1/d of the block slots, so that
                    // d consecutive launches are resident at once (HP_PIPE_OVERLAP_PREV | HP_PIPE_DEPTH(d))
    PeerLink link;  // world > 1: the last block sums the partial vector over the ranks itself (NVLink peer memory)
    int defer;      // world > 1: deferred exchange (HP_PIPE_DEFER_EXCHANGE): this step only sends, and completes the previous one
    int strict;     // serialised launch: wait for the previous grid BEFORE the first global read (the launch carries the
                    // programmatic attribute only so that block scheduling and the prologue hide the launch gap)
    int cert;       // block sums through self-certifying accumulators (see below); 0: the atomics + fence + ticket epilogue
    unsigned long long* certs;  // [kCertReplicas][kCertStride] 64-bit accumulators in the workspace (zero between launches)
    unsigned long long* trace;  // nullable profiling buffer (hp_debug_pipeline_trace): per block a header
                                // {globaltimer, clock64} at entry and exit, per warp and map 4 clock64 stamps
};

// trace buffer layout (uint64): block b at b * kTraceBlockWords: [0] globaltimer in, [1] clock in, [2] globaltimer out,
// [3] clock out, [4] SM id, [5] clock when the block's last warp left the map loop, [6..7] spare, then per warp w
// (< 16) and map jj (< kTraceMaps): 4 stamps {wait begins, data landed, refill issued, map closed}
constexpr int kTraceMaps = 8;
constexpr int kTraceHdr = 8;
constexpr int kTraceBlockWords = kTraceHdr + 16 * kTraceMaps * 4;
constexpr int kTraceMaxBlocksPerSM = 4;
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// ---- per-lane accumulators of one map ------------------------------------------------------------------
struct BulkAcc {
    float M;      // maximum so far (all lanes agree)
    int idx;      // its first flat index
    float2 s2;    // per-lane partial sum exp(p - M)
    float2 sp2;   // per-lane partial sum p
    float2 spp2;  // per-lane partial sum p^2
};

// One chunk of NITC*128 elements sitting in shared memory.  Element e of the chunk is component (e & 3) of
// float4 number (e >> 2); lane l reads float4 number it*32 + l in iteration it (conflict-free LDS.128).
template <int NITC, int LOSS, bool MULTI>
__device__ __forceinline__ void bulk_chunk(BulkAcc& A, const float4* __restrict__ buf, int chunk, int lane) {
    // pass A: maximum, and per lane the first iteration that reached the lane's maximum
    float run = -INFINITY;
    int best_it = 0;
#pragma unroll 8
    for (int it = 0; it < NITC; ++it) {
        const float4 v = buf[it * 32 + lane];
        const float t = fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w));
        best_it = (t > run) ? it : best_it;  // strict: the earlier iteration keeps ties
        run = fmaxf(run, t);
    }
    const float cm = warp_max_f32(run);
    if (!MULTI || chunk == 0 || cm > A.M) {
        // first flat index of cm: smallest (iteration, lane) among the lanes that hold it, then the component
        const unsigned key = (run == cm) ? static_cast<unsigned>(best_it * 32 + lane) : 0x7fffffffu;
        const unsigned kmin = __reduce_min_sync(0xffffffffu, key);
        const float4 v = buf[kmin];
        const int comp = (v.x == cm) ? 0 : ((v.y == cm) ? 1 : ((v.z == cm) ? 2 : 3));
        A.idx = chunk * (NITC * 128) + static_cast<int>(kmin) * 4 + comp;
        if (MULTI && (LOSS & HP_LOSS_KL)) {  // re-base the sums collected so far on the new maximum
            const float scale = (A.M == -INFINITY) ? 0.0f : ex2_approx((A.M - cm) * kLog2e);
            A.s2.x *= scale;
            A.s2.y *= scale;
        }
        A.M = cm;
    }
    // pass B: the sums, against the true maximum
    const float ms = (A.M == -INFINITY) ? 0.0f : A.M;
    const float2 l2 = make_float2(kLog2e, kLog2e), mb2 = make_float2(-ms * kLog2e, -ms * kLog2e);
    float2 s2 = A.s2, sp2 = A.sp2, spp2 = A.spp2;
#pragma unroll 8
    for (int it = 0; it < NITC; ++it) {
        const float4 v = buf[it * 32 + lane];
        const float2 lo = make_float2(v.x, v.y), hi = make_float2(v.z, v.w);
        if (LOSS & HP_LOSS_KL) {
            const float2 a0 = __ffma2_rn(lo, l2, mb2);
            const float2 a1 = __ffma2_rn(hi, l2, mb2);
            const float2 e0 = make_float2(ex2_approx(a0.x), ex2_approx(a0.y));
            const float2 e1 = make_float2(ex2_approx(a1.x), ex2_approx(a1.y));
            s2 = __fadd2_rn(s2, __fadd2_rn(e0, e1));
        }
        sp2 = __fadd2_rn(sp2, __fadd2_rn(lo, hi));
        if (LOSS & HP_LOSS_MSE) {
            spp2 = __ffma2_rn(lo, lo, spp2);
            spp2 = __ffma2_rn(hi, hi, spp2);
        }
    }
    A.s2 = s2;
    A.sp2 = sp2;
    A.spp2 = spp2;
}

// per-map losses from the reduced sums, float32 closure (a per-map loss only has float32 accuracy anyway; the
// exact part - the accumulation over maps - is the 64-bit fixed-point sum).  Same algebra as pipe_losses.
template <int LOSS>
__device__ __forceinline__ void bulk_losses(const PipeArgs& a, Centre c, float weight, float inv_hw, float vmax,
                                            float sum_exp, float sum_p, float sum_pp, const PatchSums& ps, float& mse,
                                            float& kl) {
    mse = 0.0f;
    kl = 0.0f;
    if (LOSS & HP_LOSS_MSE)  // mean over HW of 0.5*w*(p-t)^2 (loss.py:59-65); sum (p-t)^2 = sum p^2 + sum_patch t(t-2p)
        mse = 0.5f * weight * ((sum_pp + ps.e) * inv_hw);
    if (LOSS & HP_LOSS_KL) {
        const float n_bg = static_cast<float>(a.HW - pipe_patch_area(a, c));
        const float Su = fmaf(a.eps, n_bg, ps.u);
        const float Sup = fmaf(a.eps, sum_p - ps.p, ps.up);
        const float Sulogu = fmaf(n_bg, a.eps_log_eps, ps.ulogu);
        const float lse = fmaf(lg2_approx(sum_exp), kLn2, vmax);
        // Su == 0 (eps 0 and nothing pasted) -> 0/0 = NaN, like the reference
        const float L = __fdividef(Sulogu - Sup, Su) - lg2_approx(Su) * kLn2 + lse;
        kl = L * weight;
    }
}
// fixed-point accumulate of a float32 per-map loss: v * 2^40 is an exact scaling, the conversion is exact
__device__ __forceinline__ void warp_loss_add_f32(WarpLoss* w, int which, float v) {
    if (fabsf(v) < static_cast<float>(kFxLimit)) w->fx[which] += __float2ll_rn(v * 1099511627776.0f);
    else warp_loss_add_nonfinite(w, which, static_cast<double>(v));
}

// shared memory of the kernel besides the stages
constexpr int kBulkOutCap = 64;  // per-map outputs of a block that wait in shared memory for the previous grid
struct BulkShared {
    PatchSlot patch[kTileMaxPatch * 32];
    int counts[2 * HP_MAX_K];       // PCK hits / valid of this block
    unsigned long long acc[8];      // loss sums (fixed point) and non-finite counters of this block
    float4 out[kBulkOutCap];        // {x, y, maxval, weight} of the block's first maps
    long long pub[4 + 2 * HP_MAX_K + 6];
    double pub_acc[HP_MAX_K];
};

// ---- block sums through self-certifying accumulators ---------------------------------------------------------------
// The classic "REDs into a workspace, __threadfence, atomic ticket, last block reads the workspace back" epilogue puts
// three dependent L2 round trips between the last map and the end of the launch (the fence waits for the REDs, the
// ticket for the fence, the read-back for the ticket): 2-4.5 us of a ~20 us launch (profiles/r1_trace_serial.json).
// Here every accumulator certifies ITSELF: it is a 64-bit word {contributors:32 | value:32}, and every block adds
// (1 << 32 | its value) to EVERY accumulator with one fire-and-forget 64-bit RED - zeros included.  A word whose high
// half equals the grid size is complete, whatever order the REDs arrived in: no fence, no ticket, no read-back after a
// flag.  The publisher (the last block index: it owns the fewest maps) polls the 2K + 14 words with one warp - a single
// L2 round trip per poll -, zeroes them for the next launch and finalises.  One hop + one poll on the critical path.
// Layout (value field = two 16-bit halves where noted; every count is <= n_maps <= 65535, checked on the host):
// [0, K) {valid[k] : hits[k]}, [K, K+3) the six non-finite loss counters two per word, then the two 64-bit fixed-point
// loss sums as 4 + 4 limbs of 16 bits (a limb summed over <= 65535 blocks fits the 32-bit value field; sum_i limb_i << 16 i
// reproduces the two's-complement total mod 2^64).  K = 21: 32 words - one per lane of the polling warp.
// Same-address REDs serialise in the L2 slice that owns the word, so the accumulators exist in kCertReplicas copies
// (block b adds to copy b % kCertReplicas) and the publisher sums the copies: a word is complete when the contributor
// counts of its copies add up to the grid size.
constexpr size_t kCertOffsetBytes = 2048;   // of the accumulators inside the workspace (after the Workspace header)
constexpr int kCertReplicas = 4;
constexpr int kCertStride = 80;             // words per copy (>= HP_MAX_K + 11 = 75), 640 B: copies start on new lines
__device__ __forceinline__ int cert_words(int K) { return K + 3 + 8; }
__device__ __forceinline__ unsigned long long cert_load(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// last block, ONE warp: workspace -> partial (= or +=), workspace back to zero, optional finalise
// (same results as pipeline_publish, without block barriers)
// When the batch is sharded over GPUs (link.world > 1) the same warp then performs the path's one collective
// in place: it stores the vector into every rank's mailbox over NVLink as step-tagged 8-byte entries, polls
// (bounded) for the other ranks' vectors of the same step and sums them in rank order - compute and collective
// in ONE kernel, no NCCL launch and no second kernel on the step path.  All entries are integers, so every rank
// ends with bit-identical totals.
// `split_done` != ~0: the pending step is being completed by another warp of this block (peer_pending_warp); this warp
// meets it at named barrier 1 (64 threads) and only sends (peer_send_warp) with the step counter it was given.
__device__ __forceinline__ void bulk_publish_warp(const BulkArgs& t, BulkShared& sh, long long* scratch, int lane,
                                                  unsigned long long split_done = ~0ull) {
    const PipeArgs& a = t.p;
    const PeerLink& link = t.link;
    const int K = a.K, n = 4 + 2 * K + 6;
    const bool add = a.accumulate != 0;
    const bool exchange = link.world > 1;
    const bool from_cert = t.cert != 0;  // raw totals already in sh.pub (cert_collect_warp)
    for (int i = lane; i < n; i += 32) {
        long long v;
        if (i == 2) {
            v = a.n_maps;
        } else if (i == 3) {
            v = static_cast<long long>(a.n_maps) * a.HW;
        } else if (from_cert) {
            v = sh.pub[i];
        } else if (i < 2) {
            v = static_cast<long long>(*reinterpret_cast<volatile unsigned long long*>(&a.ws->acc[i]));
            a.ws->acc[i] = 0;
        } else if (i < 4 + 2 * K) {
            v = *reinterpret_cast<volatile int*>(&a.ws->counts[i - 4]);
            a.ws->counts[i - 4] = 0;
        } else {
            const int j = 2 + (i - 4 - 2 * K);
            v = static_cast<long long>(*reinterpret_cast<volatile unsigned long long*>(&a.ws->acc[j]));
            a.ws->acc[j] = 0;
        }
        if (add) v += a.partial[i];
        if (!exchange) a.partial[i] = v;
        sh.pub[i] = v;
    }
    __syncwarp();
    if (exchange) {
        // sh.pub = this rank's vector: send it, and collect / finalise this step (synchronous) or the previous one (deferred)
        if (split_done != ~0ull) {
            named_barrier<1, 64>();  // the other warp has read (and completed) the pending record
            peer_send_warp(link, a.ws, sh.pub, scratch, sh.pub_acc, K, a.partial, a.result, t.defer, split_done, lane);
        } else {
            peer_step_warp(link, a.ws, sh.pub, scratch, sh.pub_acc, K, a.partial, a.result, t.defer, lane);
        }
    } else if (a.result) {
        warp_result_from_partial(sh.pub, K, a.result, sh.pub_acc, lane);
    }
    if (lane == 0 && !from_cert) a.ws->counter = 0;
}

// NITC: iterations (of 128 elements) per chunk;  MULTI: maps span several chunks;
// W warps per block, KST private stages per warp, BPS blocks per SM (W*BPS warps and W*KST*BPS stages per SM:
// several small blocks per SM let the blocks of the NEXT launch take over an SM piecewise while this launch drains).
template <int NITC, int LOSS, bool MULTI, int W, int KST, int BPS>
__global__ void __launch_bounds__(32 * W, BPS) pipeline_bulk_kernel(const BulkArgs t) {
    extern __shared__ __align__(128) unsigned char s_dyn[];
    __shared__ BulkShared sh;
    __shared__ WarpLoss s_wl[W];
    constexpr int kChunkBytes = NITC * 512;
    constexpr int kChunkElems = NITC * 128;
    const PipeArgs& a = t.p;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_chunks = MULTI ? t.n_chunks : 1;
    // maps of this block: blockIdx.x + j*gridDim.x ; of this warp: j = warp + jj*W
    const int n_local = (a.n_maps - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
    const int n_mine = (n_local > warp) ? (n_local - warp + W - 1) / W : 0;
    const int q_total = n_mine * n_chunks;  // chunk loads of this warp, in order q = jj*n_chunks + c

    unsigned char* my_stage = s_dyn + static_cast<size_t>(warp) * KST * kChunkBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_dyn + static_cast<size_t>(W) * KST * kChunkBytes) + warp * KST;
    const uint32_t stage_u32 = smem_addr(my_stage), bar_u32 = smem_addr(bars);
    const size_t map_stride = static_cast<size_t>(gridDim.x) * W * a.HW;  // between consecutive maps of this warp
    const float* my_first = a.pred + (static_cast<size_t>(blockIdx.x) + static_cast<size_t>(warp) * gridDim.x) * a.HW;

    // Programmatic dependent launch (no-ops unless the launch carries the attribute): the next launch on the
    // stream may start filling SMs as soon as this grid's blocks retire; nothing of a launch is WRITTEN to
    // global memory (and the shared workspace is not touched) before griddep_wait() has seen the previous grid
    // complete - per-map outputs wait in shared memory until then.
    griddep_launch_dependents();
    // serialised launch: nothing of this launch is read or written before the previous grid has completed and flushed;
    // only block scheduling and the barrier set-up above the first global access overlap its tail
    if (t.strict) griddep_wait();
    unsigned long long* trace = t.trace ? t.trace + static_cast<size_t>(blockIdx.x) * kTraceBlockWords : nullptr;
    if (trace && threadIdx.x == 0) {
        trace[0] = global_timer_ns();
        trace[1] = static_cast<unsigned long long>(clock64());
        unsigned int smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        trace[4] = smid;
    }
    unsigned long long* wtrace = (trace && lane == 0 && warp < 16) ? trace + kTraceHdr + warp * (kTraceMaps * 4) : nullptr;
    // ---- prologue ----------------------------------------------------------------------------------------------
    // Order matters: (1) arm the ring (the init fence would otherwise wait for the loads below), (2) ISSUE the
    // small global loads, (3) request the bulk copies, (4) consume the small loads.  Once 148 x 192 KB of bulk
    // traffic is queued a plain load waits microseconds behind it (the first version of this kernel loaded each
    // map's keypoint right before use and spent a third of its time there).
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < KST; ++s) mbar_init(bar_u32 + 8 * s, 1);
        mbar_init_fence();
    }
    // keypoints are lane-distributed: lane l of a warp holds the keypoint of the warp's map 32*batch + l
    const bool joints16 = (reinterpret_cast<uintptr_t>(a.joints) & 15u) == 0;
    double jx_cur = 0.0, jy_cur = 0.0, jx_nxt = 0.0, jy_nxt = 0.0;
    float vis_cur = 0.0f, vis_nxt = 0.0f;
    auto load_keypoints = [&](int batch, double& jx, double& jy, float& vis) {
        const int jj = batch * 32 + lane;
        if (jj < n_mine) {
            const int m = static_cast<int>(blockIdx.x) + (warp + jj * W) * static_cast<int>(gridDim.x);
            const double* jp = a.joints + 2 * static_cast<size_t>(m);
            if (joints16) {
                const double2 j2 = *reinterpret_cast<const double2*>(jp);
                jx = j2.x;
                jy = j2.y;
            } else {
                jx = jp[0];
                jy = jp[1];
            }
            vis = a.vis[m];
        }
    };
    load_keypoints(0, jx_cur, jy_cur, vis_cur);
    const int side = 2 * a.tmp + 1, n_patch = side * side;
    // per-lane patch table: slot k of lane l is patch pixel l + 32k (row-major); thread t fills slots t, t + 32W, ...
    constexpr int kSlotsPerThread = (kTileMaxPatch * 32 + 32 * W - 1) / (32 * W);
    PatchSlot my_slot[kSlotsPerThread];
#pragma unroll
    for (int r = 0; r < kSlotsPerThread; ++r) {
        const int i = static_cast<int>(threadIdx.x) + r * 32 * W;
        my_slot[r].dx = 1 << 20;
        my_slot[r].dy = 0;
        my_slot[r].t = 0.f;
        my_slot[r].ulogu = 0.f;
        if (i < n_patch) {
            uint32_t ry, rx;
            a.sdiv.divmod(static_cast<uint32_t>(i), ry, rx);
            my_slot[r].dx = static_cast<int>(rx) - a.tmp;
            my_slot[r].dy = static_cast<int>(ry) - a.tmp;
            my_slot[r].t = a.tab[my_slot[r].dx * my_slot[r].dx + my_slot[r].dy * my_slot[r].dy];
        }
    }
    // request the first KST chunks
    const uint64_t pol = l2_evict_first_policy();
    int q_load = 0, load_jj = 0, load_c = 0;  // next chunk to request
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < KST; ++s) {
            if (q_load < q_total) {
                mbar_arrive_expect_tx(bar_u32 + 8 * s, kChunkBytes);
                bulk_load(stage_u32 + s * kChunkBytes, my_first + load_jj * map_stride + load_c * kChunkElems, kChunkBytes,
                          bar_u32 + 8 * s, pol);
                ++q_load;
                if (++load_c == n_chunks) {
                    load_c = 0;
                    ++load_jj;
                }
            }
        }
    }
    load_keypoints(1, jx_nxt, jy_nxt, vis_nxt);
    // patch table and block accumulators
#pragma unroll
    for (int r = 0; r < kSlotsPerThread; ++r) {
        const int i = static_cast<int>(threadIdx.x) + r * 32 * W;
        if (i < kTileMaxPatch * 32) {
            if (i < n_patch) {
                const float u = my_slot[r].t + a.eps;
                my_slot[r].ulogu = (u != 0.0f) ? u * logf(u) : 0.0f;
            }
            sh.patch[i] = my_slot[r];
        }
    }
    for (int i = threadIdx.x; i < 2 * HP_MAX_K; i += blockDim.x) sh.counts[i] = 0;
    if (threadIdx.x < 8) sh.acc[threadIdx.x] = 0;
    if (lane == 0) {
        s_wl[warp].fx[0] = s_wl[warp].fx[1] = 0;
        for (int i = 0; i < 6; ++i) s_wl[warp].cls[i] = 0;
    }
    __syncthreads();  // the only block barrier before the epilogue

    WarpLoss* wl = &s_wl[warp];
    const float inv_hw = 1.0f / static_cast<float>(a.HW);
    bool dep_ok = t.strict != 0;  // griddep_wait() already executed by this warp
    int q = 0;            // chunk being consumed
    for (int jj = 0; jj < n_mine; ++jj) {
        const int j = warp + jj * W;
        const int map = static_cast<int>(blockIdx.x) + j * static_cast<int>(gridDim.x);
        // ---- everything that does not need the prediction, while the chunk is in flight ----------------------
        if (jj != 0 && (jj & 31) == 0) {  // next batch of 32 keypoints: requested 32 maps ago
            jx_cur = jx_nxt;
            jy_cur = jy_nxt;
            vis_cur = vis_nxt;
            load_keypoints((jj >> 5) + 1, jx_nxt, jy_nxt, vis_nxt);
        }
        float weight;
        const Centre c = pipe_centre(a, __shfl_sync(0xffffffffu, jx_cur, jj & 31), __shfl_sync(0xffffffffu, jy_cur, jj & 31),
                                     __shfl_sync(0xffffffffu, vis_cur, jj & 31), weight);
        const bool pasted = c.y != kNoPaste;
        PatchSums ps{0.f, 0.f, 0.f, 0.f, 0.f};
        int off[kTileMaxPatch];
        float tv[kTileMaxPatch];
        if (LOSS != 0) {
#pragma unroll
            for (int k = 0; k < kTileMaxPatch; ++k) {
                const PatchSlot sl = sh.patch[k * 32 + lane];
                const int x = c.x + sl.dx, y = c.y + sl.dy;
                const bool in = pasted && static_cast<unsigned>(x) < static_cast<unsigned>(a.W) &&
                                static_cast<unsigned>(y) < static_cast<unsigned>(a.H);
                off[k] = in ? y * a.W + x : -1;
                tv[k] = in ? sl.t : 0.0f;
                if (LOSS & HP_LOSS_KL) {
                    ps.ulogu += in ? sl.ulogu : 0.0f;
                    ps.u += (tv[k] != 0.0f) ? tv[k] + a.eps : 0.0f;
                }
            }
            if (LOSS & HP_LOSS_KL) {
                const float r = warp_sum3_scattered(ps.u, ps.ulogu, 0.0f, lane);
                ps.u = __shfl_sync(0xffffffffu, r, 0);
                ps.ulogu = __shfl_sync(0xffffffffu, r, 8);
            }
        }
        BulkAcc A;
        A.M = -INFINITY;
        A.idx = 0;
        A.s2 = A.sp2 = A.spp2 = make_float2(0.f, 0.f);
        // ---- the map, chunk by chunk, from shared memory ----------------------------------------------------------
        for (int ch = 0; ch < n_chunks; ++ch, ++q) {
            const int s = q % KST;
            const uint32_t parity = static_cast<uint32_t>(q / KST) & 1u;
            const float* buf = reinterpret_cast<const float*>(my_stage + s * kChunkBytes);
            if (wtrace && jj < kTraceMaps && ch == 0) wtrace[jj * 4 + 0] = static_cast<unsigned long long>(clock64());
            mbar_wait(bar_u32 + 8 * s, parity);
            if (wtrace && jj < kTraceMaps && ch == 0) wtrace[jj * 4 + 1] = static_cast<unsigned long long>(clock64());
            bulk_chunk<NITC, LOSS, MULTI>(A, reinterpret_cast<const float4*>(buf), ch, lane);
            if (LOSS != 0) {
#pragma unroll
                for (int k = 0; k < kTileMaxPatch; ++k) {
                    const int rel = off[k] - ch * kChunkElems;
                    const bool here = MULTI ? (static_cast<unsigned>(rel) < static_cast<unsigned>(kChunkElems)) : (rel >= 0);
                    const float pk = here ? buf[rel] : 0.0f;
                    const float tk = here ? tv[k] : 0.0f;
                    if (LOSS & HP_LOSS_KL) {
                        const float u = (tk != 0.0f) ? tk + a.eps : 0.0f;
                        ps.up = fmaf(u, pk, ps.up);
                        ps.p += pk;
                    }
                    if (LOSS & HP_LOSS_MSE) ps.e = fmaf(tk, tk - 2.0f * pk, ps.e);
                }
            }
            // the stage has been read out (every lane's values are in registers): request the chunk KST ahead
            __syncwarp();
            if (lane == 0 && q_load < q_total) {
                mbar_arrive_expect_tx(bar_u32 + 8 * s, kChunkBytes);
                bulk_load(stage_u32 + s * kChunkBytes, my_first + load_jj * map_stride + load_c * kChunkElems, kChunkBytes,
                          bar_u32 + 8 * s, pol);
                ++q_load;
                if (++load_c == n_chunks) {
                    load_c = 0;
                    ++load_jj;
                }
            }
            if (wtrace && jj < kTraceMaps && ch == n_chunks - 1) wtrace[jj * 4 + 2] = static_cast<unsigned long long>(clock64());
        }
        // ---- warp reductions ---------------------------------------------------------------------------------------
        float sum_exp, sum_p, sum_pp;
        {
            const float r = warp_sum3_scattered(A.s2.x + A.s2.y, A.sp2.x + A.sp2.y, A.spp2.x + A.spp2.y, lane);
            sum_exp = __shfl_sync(0xffffffffu, r, 0);
            sum_p = __shfl_sync(0xffffffffu, r, 8);
            sum_pp = __shfl_sync(0xffffffffu, r, 16);
        }
        if (LOSS != 0) {
            const float r = warp_sum3_scattered(ps.up, ps.p, ps.e, lane);
            ps.up = __shfl_sync(0xffffffffu, r, 0);
            ps.p = __shfl_sync(0xffffffffu, r, 8);
            ps.e = __shfl_sync(0xffffffffu, r, 16);
        }
        ArgMax am;
        am.v = A.M;
        am.i = A.idx;
        if (sum_p != sum_p) {
            // a NaN (or +inf with -inf) is in the map: redo the argmax with numpy's exact rules from memory (L2)
            ArgMax sx = am_init();
            const float4* m4 = reinterpret_cast<const float4*>(a.pred + static_cast<size_t>(map) * a.HW);
            for (int e4 = lane; e4 < a.HW / 4; e4 += 32) am_scan4<true>(sx, ldg_stream4(m4 + e4), e4 * 4);
            am = warp_argmax(sx, lane);
            sum_exp = __int_as_float(0x7fc00000);  // log_softmax of a map holding a NaN is NaN
        }
        // ---- closure: decode, PCK, losses, results -----------------------------------------------------------------
        uint32_t qy, qx;
        a.wdiv.divmod(static_cast<uint32_t>(am.i), qy, qx);
        const float keep = (am.v > 0.0f) ? 1.0f : 0.0f;  // NaN -> 0 (keypoint_detection.py:31-34)
        const float px = static_cast<float>(qx) * keep, py = static_cast<float>(qy) * keep;
        // decoding the generated target: its unique maximum (exactly 1.0) sits on the centre when pasted;
        // the all-zero map decodes to the masked (0,0)   
        const float tx = pasted ? static_cast<float>(c.x) : 0.0f, ty = pasted ? static_cast<float>(c.y) : 0.0f;
        int valid, hit;
        pipe_pck(a, px, py, tx, ty, valid, hit);
        float mse, kl;
        bulk_losses<LOSS>(a, c, weight, inv_hw, am.v, sum_exp, sum_p, sum_pp, ps, mse, kl);
        if (j >= kBulkOutCap && !dep_ok) {  // results beyond the shared-memory buffer go straight to global memory
            griddep_wait();
            dep_ok = true;
        }
        if (lane == 0) {
            if (LOSS & HP_LOSS_MSE) warp_loss_add_f32(wl, 0, mse);
            if (LOSS & HP_LOSS_KL) warp_loss_add_f32(wl, 1, kl);
            if (j < kBulkOutCap) {
                sh.out[j] = make_float4(px, py, am.v, weight);
            } else {
                *reinterpret_cast<float2*>(a.pred_xy + 2 * static_cast<size_t>(map)) = make_float2(px, py);
                if (a.maxvals) a.maxvals[map] = am.v;
                if (a.weight_out) a.weight_out[map] = weight;
            }
            const int k = map - static_cast<int>(t.kdiv.div(static_cast<uint32_t>(map))) * a.K;
            if (valid) atomicAdd(&sh.counts[a.K + k], 1);
            if (hit) atomicAdd(&sh.counts[k], 1);
        }
        if (wtrace && jj < kTraceMaps) wtrace[jj * 4 + 3] = static_cast<unsigned long long>(clock64());
    }

    // ---- epilogue: block sums -> workspace (integer atomics: exact, order-free), last block publishes ---------
    if (lane == 0) {
        for (int w = 0; w < 2; ++w)
            if (wl->fx[w] != 0) atomicAdd(&sh.acc[w], static_cast<unsigned long long>(wl->fx[w]));
        for (int i = 0; i < 6; ++i)
            if (wl->cls[i] != 0) atomicAdd(&sh.acc[2 + i], static_cast<unsigned long long>(wl->cls[i]));
    }
    __syncthreads();
    if (trace && threadIdx.x == 0) trace[5] = static_cast<unsigned long long>(clock64());
    if (!dep_ok) griddep_wait();  // (warps without maps) previous grid complete: outputs and workspace may be written
    if (t.cert != 0) {
        // ---- self-certifying accumulators: one fire-and-forget RED per accumulator and block, the publisher polls ------
        const bool publisher = blockIdx.x == gridDim.x - 1;
        const int K = a.K, nw = cert_words(K);
        if (warp == 0) {
            for (int i = lane; i < nw; i += 32) {
                unsigned int v;
                if (i < K) {
                    v = static_cast<unsigned int>(sh.counts[i]) | (static_cast<unsigned int>(sh.counts[K + i]) << 16);
                } else if (i < K + 3) {
                    const int c = 2 * (i - K);
                    v = static_cast<unsigned int>(sh.acc[2 + c]) | (static_cast<unsigned int>(sh.acc[3 + c]) << 16);
                } else {
                    const int f = i - K - 3;  // limb f & 3 of loss sum f >> 2
                    v = static_cast<unsigned int>((sh.acc[f >> 2] >> (16 * (f & 3))) & 0xffffull);
                }
                atomicAdd(&t.certs[(blockIdx.x % kCertReplicas) * kCertStride + i],
                          (1ull << 32) | static_cast<unsigned long long>(v));  // RED.E.ADD.64, no return
            }
        }
        const bool exchange = t.link.world > 1;
        if (!(publisher && warp == 0)) {  // the other warps deliver the buffered per-map outputs meanwhile
            const int n_buf = n_local < kBulkOutCap ? n_local : kBulkOutCap;
            const int first = publisher ? static_cast<int>(threadIdx.x) - 32 : static_cast<int>(threadIdx.x);
            const int step = publisher ? 32 * (W - 1) : 32 * W;
            for (int jb = first; jb < n_buf; jb += step) {
                const int map = static_cast<int>(blockIdx.x) + jb * static_cast<int>(gridDim.x);
                const float4 o = sh.out[jb];
                *reinterpret_cast<float2*>(a.pred_xy + 2 * static_cast<size_t>(map)) = make_float2(o.x, o.y);
                if (a.maxvals) a.maxvals[map] = o.z;
                if (a.weight_out) a.weight_out[map] = o.w;
            }
            if (publisher && warp == 1 && exchange) {
                // the step that is still outstanding from the previous launch: collected and finalised by THIS warp while
                // warp 0 waits for the blocks' sums (scratch: the idle stages behind warp 0's area)
                long long* scratch1 = reinterpret_cast<long long*>(s_dyn) + kPeerScratchWords + HP_MAX_K;
                peer_pending_warp(t.link, a.ws, scratch1, reinterpret_cast<double*>(scratch1 + kPeerScratchWords), lane);
                named_barrier<1, 64>();
            }
        } else {
            // ONE warp: poll until every accumulator has heard from every block, take the totals, zero the words
            unsigned long long steps_done = ~0ull;
            if (exchange) steps_done = peer_load(peer_counter(t.link.mailbox[t.link.rank], t.link.world));  // final: the previous grid is complete
            const unsigned int want = gridDim.x;
            const long long t0 = clock64();
            unsigned long long w0 = 0ull, w1 = 0ull, w2 = 0ull;  // words lane, lane + 32, lane + 64 (K + 11 <= 75)
            bool done = false;
            auto word = [&](int i) {  // sum of the copies of word i (all kCertReplicas loads in flight)
                unsigned long long c[kCertReplicas];
#pragma unroll
                for (int r = 0; r < kCertReplicas; ++r) c[r] = cert_load(t.certs + r * kCertStride + i);
                unsigned long long tot = 0ull;
#pragma unroll
                for (int r = 0; r < kCertReplicas; ++r) tot += c[r];
                return tot;
            };
            while (!done) {
                bool ok = true;
                if (lane < nw) { w0 = word(lane); ok &= static_cast<unsigned int>(w0 >> 32) == want; }
                if (lane + 32 < nw) { w1 = word(lane + 32); ok &= static_cast<unsigned int>(w1 >> 32) == want; }
                if (lane + 64 < nw) { w2 = word(lane + 64); ok &= static_cast<unsigned int>(w2 >> 32) == want; }
                done = __all_sync(0xffffffffu, ok);
                if (!done && clock64() - t0 > 4000000000ll) __trap();  // ~2 s: a block of this grid never reported (it faulted)
            }
            // raw partial vector in sh.pub: [0,1] loss sums, [4, 4+2K) counts, [4+2K, 4+2K+6) non-finite counters
            // (the 16-bit halves of a word were summed over the blocks without carries: each total is <= n_maps <= 65535)
            unsigned long long* limb = reinterpret_cast<unsigned long long*>(sh.pub_acc);  // 8 words of scratch
            auto take = [&](int i, unsigned long long w) {
                if (i >= nw) return;
                const unsigned long long v = w & 0xffffffffull;
                if (i < K) {
                    sh.pub[4 + i] = static_cast<long long>(v & 0xffffull);
                    sh.pub[4 + K + i] = static_cast<long long>(v >> 16);
                } else if (i < K + 3) {
                    sh.pub[4 + 2 * K + 2 * (i - K)] = static_cast<long long>(v & 0xffffull);
                    sh.pub[4 + 2 * K + 2 * (i - K) + 1] = static_cast<long long>(v >> 16);
                } else {
                    limb[i - K - 3] = v;
                }
#pragma unroll
                for (int r = 0; r < kCertReplicas; ++r)  // ready for the next launch (ordered before it by the grid's completion)
                    t.certs[r * kCertStride + i] = 0ull;
            };
            take(lane, w0); take(lane + 32, w1); take(lane + 64, w2);
            __syncwarp();
            if (lane < 2) {
                unsigned long long tot = 0ull;
                for (int i = 0; i < 4; ++i) tot += limb[4 * lane + i] << (16 * i);
                sh.pub[lane] = static_cast<long long>(tot);
            }
            __syncwarp();
            // (the block's stages are idle by now and serve as the exchange's scratch space)
            bulk_publish_warp(t, sh, reinterpret_cast<long long*>(s_dyn), lane, steps_done);
        }
    } else if (warp == 0) {
        for (int i = lane; i < 8 + 2 * a.K; i += 32) {
            if (i < 8) {
                const unsigned long long v = sh.acc[i];
                if (v != 0) atomicAdd(&a.ws->acc[i], v);
            } else {
                const int v = sh.counts[i - 8];
                if (v != 0) atomicAdd(&a.ws->counts[i - 8], v);
            }
        }
        __threadfence();  // every lane: its own atomics before the ticket
        __syncwarp();
        int last = 0;
        if (lane == 0) {
            last = (atomicAdd(&a.ws->counter, 1u) == gridDim.x - 1) ? 1 : 0;
            if (last) __threadfence();
        }
        last = __shfl_sync(0xffffffffu, last, 0);
        // (the block's stages are idle by now - every copy was consumed before the barrier above - and serve as the
        // exchange's scratch space)
        if (last) bulk_publish_warp(t, sh, reinterpret_cast<long long*>(s_dyn), lane);
    } else {
        // the other warps deliver the buffered per-map outputs meanwhile
        const int n_buf = n_local < kBulkOutCap ? n_local : kBulkOutCap;
        for (int jb = static_cast<int>(threadIdx.x) - 32; jb < n_buf; jb += 32 * (W - 1)) {
            const int map = static_cast<int>(blockIdx.x) + jb * static_cast<int>(gridDim.x);
            const float4 o = sh.out[jb];
            *reinterpret_cast<float2*>(a.pred_xy + 2 * static_cast<size_t>(map)) = make_float2(o.x, o.y);
            if (a.maxvals) a.maxvals[map] = o.z;
            if (a.weight_out) a.weight_out[map] = o.w;
        }
    }
    if (trace && threadIdx.x == 0) {
        trace[2] = global_timer_ns();
        trace[3] = static_cast<unsigned long long>(clock64());
    }
}

}  // namespace hp
