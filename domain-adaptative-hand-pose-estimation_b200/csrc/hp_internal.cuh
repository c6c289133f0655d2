// hp_internal.cuh - launch helpers shared between translation units of libhp_b200.so.
#pragma once
#include "hp_common.cuh"

namespace hp {

// decode n_maps maps; any of preds/maxvals/idx/centres may be null.
// centres[map] = (int(px) >> shift, int(py) >> shift)  -- regda_7.py:3033 / :3195 `(preds / d).astype(int)`
int launch_decode(const float* heat, int n_maps, int H, int W, float* preds, float* maxvals, int32_t* idx,
                  int32_t* centres, int shift, cudaStream_t stream);

// ---- peer mailboxes (hp_peer.cu) ---------------------------------------------------------------------------------
// Layout (uint64 entries): slots[2 parities][world sources][kPeerSlotEntries], then 8 entries (entry 0: step counter).
// Protocol ("low latency", like NCCL's LL): every 8-byte entry carries 32 bits of payload and the 32-bit step
// tag, and an aligned 8-byte store is single-copy atomic - so the sender just fires its stores (no fence, no
// separate flag, one NVLink hop) and the receiver polls each entry until its tag is the current step.  An int64
// of the partial vector travels as two entries.  Parity double-buffering: a rank can be at most one step ahead of
// the slowest rank (it cannot finish step s+1 before it has received everybody's step-(s+1) vector).
constexpr int kPeerSlotEntries = 128;   // >= 2 * (4 + 2K + 6)  ->  K <= 27
constexpr int kPeerMaxWorld = 16;
struct PeerLink {                       // passed by value to kernels that do the exchange themselves
    unsigned long long* mailbox[kPeerMaxWorld];  // base of every rank's mailbox as mapped in this process
    int rank, world;                             // world <= 1: no exchange
};
__device__ __forceinline__ unsigned long long* peer_slot(unsigned long long* base, int world, int parity, int src) {
    return base + (static_cast<size_t>(parity) * world + src) * kPeerSlotEntries;
}
__device__ __forceinline__ unsigned long long* peer_counter(unsigned long long* base, int world) {
    return base + static_cast<size_t>(2) * world * kPeerSlotEntries;
}
__device__ __forceinline__ void peer_store(unsigned long long* p, unsigned long long v) {
    asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long peer_load(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// ONE warp: vec[0..n) (shared memory; this rank's int64 vector) -> sum over all ranks, in rank order, in place.
// The step number is counted on the device (the rank's own mailbox), so captured graphs of steps replay.
// Returns non-zero in every lane if a peer did not arrive within ~2 s.
__device__ __forceinline__ int peer_exchange_warp(const PeerLink& link, long long* vec, int n, int lane) {
    const int world = link.world, rank = link.rank;
    unsigned long long* counter = peer_counter(link.mailbox[rank], world);
    const unsigned long long seq = peer_load(counter) + 1ull;
    const unsigned long long tag = (seq & 0xffffffffull) << 32;
    const int parity = static_cast<int>(seq & 1ull);
    for (int dst = 0; dst < world; ++dst) {
        unsigned long long* slot = peer_slot(link.mailbox[dst], world, parity, rank);
        for (int w = lane; w < n; w += 32) {
            const unsigned long long v = static_cast<unsigned long long>(vec[w]);
            peer_store(slot + 2 * w, tag | (v & 0xffffffffull));
            peer_store(slot + 2 * w + 1, tag | (v >> 32));
        }
    }
    int timeout = 0;
    const long long t0 = clock64();
    for (int w = lane; w < n; w += 32) {
        long long tot = 0;
        for (int src = 0; src < world; ++src) {
            const unsigned long long* slot = peer_slot(link.mailbox[rank], world, parity, src);
            unsigned long long lo, hi;
            while (true) {
                lo = peer_load(slot + 2 * w);
                hi = peer_load(slot + 2 * w + 1);
                if ((lo & 0xffffffff00000000ull) == tag && (hi & 0xffffffff00000000ull) == tag) break;
                if (clock64() - t0 > 4000000000ll) {  // ~2 s: a peer never arrived
                    timeout = 1;
                    break;
                }
            }
            tot += static_cast<long long>((hi << 32) | (lo & 0xffffffffull));
        }
        vec[w] = tot;
    }
    timeout = __any_sync(0xffffffffu, timeout);
    if (lane == 0) peer_store(counter, seq);
    __syncwarp();
    return timeout;
}

// the per-step exchange + finalise over peer mailboxes (hp_peer.cu); `overlap` != 0: programmatic dependent launch
int launch_finalize_peer(const long long* partial, void* const* mailboxes, int rank, int world, int K, long long seq,
                         long long* partial_out, double* result, int overlap, cudaStream_t stream);

}  // namespace hp
