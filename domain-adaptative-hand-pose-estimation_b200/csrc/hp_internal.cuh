// hp_internal.cuh - launch helpers shared between translation units of libhp_b200.so.
#pragma once
#include "hp_common.cuh"

namespace hp {

// decode n_maps maps; any of preds/maxvals/idx/centres may be null.
// centres[map] = (int(px) >> shift, int(py) >> shift)  -- regda_7.py:3033 / :3195 `(preds / d).astype(int)`
int launch_decode(const float* heat, int n_maps, int H, int W, float* preds, float* maxvals, int32_t* idx,
                  int32_t* centres, int shift, cudaStream_t stream);

// ---- peer mailboxes (hp_peer.cu) ---------------------------------------------------------------------------------
// Layout (uint64 entries): slots[kPeerRing steps][world sources][kPeerSlotEntries], then 8 entries (entry 0: step counter).
// Protocol ("low latency", like NCCL's LL): every 8-byte entry carries 32 bits of payload and the 32-bit step
// tag, and an aligned 8-byte store is single-copy atomic - so the sender just fires its stores (no fence, no
// separate flag, one NVLink hop) and the receiver polls each entry until its tag is the step it waits for.  An int64
// of the partial vector travels as two entries.  Step q uses ring position q % kPeerRing.  A rank is never more than
// two steps ahead of the slowest rank's SEND (synchronous exchange: it cannot finish step q before everybody has sent
// step q; deferred exchange: step q+1 is only finished once everybody has sent step q), so a ring position is rewritten
// long after every reader has consumed it.
constexpr int kPeerSlotEntries = 128;   // >= 2 * (4 + 2K + 6)  ->  K <= 27
constexpr int kPeerRing = 16;
constexpr int kPeerMaxWorld = 16;
struct PeerLink {                       // passed by value to kernels that do the exchange themselves
    unsigned long long* mailbox[kPeerMaxWorld];  // base of every rank's mailbox as mapped in this process
    int rank, world;                             // world <= 1: no exchange
};
__device__ __forceinline__ unsigned long long* peer_slot(unsigned long long* base, int world, unsigned long long step, int src) {
    return base + (static_cast<size_t>(step % kPeerRing) * world + src) * kPeerSlotEntries;
}
__device__ __forceinline__ unsigned long long* peer_counter(unsigned long long* base, int world) {
    return base + static_cast<size_t>(kPeerRing) * world * kPeerSlotEntries;
}
__device__ __forceinline__ void peer_store2(unsigned long long* p, unsigned long long a, unsigned long long b) {
    asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(a), "l"(b) : "memory");
}
__device__ __forceinline__ void peer_store(unsigned long long* p, unsigned long long v) {
    asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ ulonglong2 peer_load2(const unsigned long long* p) {
    ulonglong2 v;
    asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long peer_load(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
constexpr int kPeerRecvWords = kPeerMaxWorld * (kPeerSlotEntries / 2);  // landing area of peer_recv_sum (world * n words)
constexpr int kPeerScratchWords = kPeerRecvWords + 144;                 // + one partial vector (4 + 2*HP_MAX_K + 6 <= 144): int64 words of shared-memory scratch
// what the exchange can carry: a K-joint partial vector is n = 4+2K+6 int64 = 2n tagged entries per source slot, and
// the receiver tracks world*n (source, word) pairs in 32 lanes x 32 pending bits.  Every entry point that accepts a
// PeerLink with world > 1 checks this (K <= 27; the slots are NOT sized from HP_MAX_K).
inline bool peer_shape_ok(int K, int world) {
    const int n = 4 + 2 * K + 6;
    return world <= 1 || (world <= kPeerMaxWorld && 2 * n <= kPeerSlotEntries && world * n <= 1024);
}
// ONE warp: this rank's vector of step `seq` -> every rank's mailbox (its own included): a word travels as two tagged
// 8-byte entries, written with one 16-byte store.  Fire and forget.
__device__ __forceinline__ void peer_send(const PeerLink& link, const long long* vec, int n, unsigned long long seq, int lane) {
    const unsigned long long tag = (seq & 0xffffffffull) << 32;
    for (int dst = 0; dst < link.world; ++dst) {
        unsigned long long* slot = peer_slot(link.mailbox[dst], link.world, seq, link.rank);
        for (int w = lane; w < n; w += 32) {
            const unsigned long long v = static_cast<unsigned long long>(vec[w]);
            peer_store2(slot + 2 * w, tag | (v & 0xffffffffull), tag | (v >> 32));
        }
    }
}
// ONE warp: wait (bounded, ~2 s) until every rank's vector of step `seq` sits in this rank's mailbox and write the sum
// over the ranks, in rank order, to out[0..n) (shared memory).  `scratch`: kPeerScratchWords int64 of shared memory.
// The (source, word) pairs are spread over the lanes and polled eight at a time, so a pass over all sources costs one
// or two memory round trips however many ranks there are (polling source after source cost a round trip per rank:
// 5.6 us at 8 GPUs).  Returns non-zero in every lane on a timeout (missing contributions count as zero).
template <int kBatch = 8>
__device__ __forceinline__ int peer_recv_sum(const PeerLink& link, long long* out, long long* scratch, int n,
                                             unsigned long long seq, int lane) {
    const int world = link.world;
    const unsigned long long tag = (seq & 0xffffffffull) << 32;
    const unsigned long long tag_mask = 0xffffffff00000000ull;
    const int total = world * n;
    const int kcount = (total + 31) >> 5;  // <= 32
    unsigned pending = 0;
    for (int k = 0; k < kcount; ++k)
        if (lane + 32 * k < total) pending |= 1u << k;
    const unsigned long long* mine = peer_slot(link.mailbox[link.rank], world, seq, 0);
    int timeout = 0;
    const long long t0 = clock64();
    while (__any_sync(0xffffffffu, pending != 0)) {
        for (int kb = 0; kb < kcount; kb += kBatch) {
            ulonglong2 v[kBatch];
#pragma unroll
            for (int j = 0; j < kBatch; ++j) {
                const int k = kb + j, p = lane + 32 * k;
                if ((pending >> k) & 1u) {
                    const int src = p / n, w = p - src * n;
                    v[j] = peer_load2(mine + static_cast<size_t>(src) * kPeerSlotEntries + 2 * w);
                }
            }
#pragma unroll
            for (int j = 0; j < kBatch; ++j) {
                const int k = kb + j, p = lane + 32 * k;
                if (((pending >> k) & 1u) && (v[j].x & tag_mask) == tag && (v[j].y & tag_mask) == tag) {
                    scratch[p] = static_cast<long long>((v[j].y << 32) | (v[j].x & 0xffffffffull));
                    pending &= ~(1u << k);
                }
            }
        }
        if (clock64() - t0 > 4000000000ll) {  // ~2 s: a peer never arrived
            timeout = pending != 0;
            for (int k = 0; k < kcount; ++k)
                if ((pending >> k) & 1u) scratch[lane + 32 * k] = 0;
            pending = 0;
        }
    }
    timeout = __any_sync(0xffffffffu, timeout);
    __syncwarp();
    for (int w = lane; w < n; w += 32) {
        long long tot = 0;
        for (int src = 0; src < world; ++src) tot += scratch[src * n + w];  // rank order
        out[w] = tot;
    }
    __syncwarp();
    return timeout;
}
// The synchronous exchange of one step: vec[0..n) (shared memory; this rank's vector) -> totals over the ranks, in
// place.  The step number is counted on the device (the rank's own mailbox), so captured graphs of steps replay.
// A timed-out step is NOT counted: the next step re-uses its number, so ranks that did complete it and ranks that did
// not cannot drift apart silently (the caller sees the poisoned result and stops).
__device__ __forceinline__ int peer_exchange_warp(const PeerLink& link, long long* vec, long long* scratch, int n, int lane) {
    unsigned long long* counter = peer_counter(link.mailbox[link.rank], link.world);
    const unsigned long long seq = peer_load(counter) + 1ull;
    peer_send(link, vec, n, seq, lane);
    const int timeout = peer_recv_sum(link, vec, scratch, n, seq, lane);
    if (lane == 0 && !timeout) peer_store(counter, seq);
    __syncwarp();
    return timeout;
}

// ---- deferred exchange (HP_PIPE_DEFER_EXCHANGE) --------------------------------------------------------------------
// In a train of sharded steps the synchronous exchange puts an NVLink round trip - and the slowest rank's jitter - on
// every step's critical path (SCALE_r01: 16.2 -> 18.7 us per step from 1 to 8 GPUs).  Deferred: step q only SENDS its
// vector; the totals of step q-1, which every rank sent a whole step ago, are collected and finalised into step q-1's
// OWN output buffers (remembered in the workspace) by step q's publisher, and the last step of a train by a one-warp
// flush kernel (hp_pipeline_flush_peer).  Nothing on the step path waits for a peer that is less than a step late.
struct PeerPending {
    long long* partial;        // output buffers of the step whose totals are still outstanding
    double* result;
    unsigned long long seq;    // its step number
    int K;
    int valid;
};
constexpr size_t kPendOffsetBytes = 1024;  // of the PeerPending record inside the workspace (zero-initialised: invalid)
__device__ __forceinline__ PeerPending* peer_pending(Workspace* ws) {
    return reinterpret_cast<PeerPending*>(reinterpret_cast<unsigned char*>(ws) + kPendOffsetBytes);
}

// the per-step exchange + finalise over peer mailboxes (hp_peer.cu); `overlap` != 0: programmatic dependent launch;
// `workspace` (nullable unless defer): where a deferred step is remembered; `defer`: see "deferred exchange" above
int launch_finalize_peer(const long long* partial, void* const* mailboxes, int rank, int world, int K, long long seq,
                         long long* partial_out, double* result, int overlap, void* workspace, int defer,
                         cudaStream_t stream);

}  // namespace hp
