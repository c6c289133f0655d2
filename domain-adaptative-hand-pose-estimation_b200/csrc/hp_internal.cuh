// hp_internal.cuh - launch helpers shared between translation units of libhp_b200.so.
#pragma once
#include "hp_common.cuh"

namespace hp {

// decode n_maps maps; any of preds/maxvals/idx/centres may be null.
// centres[map] = (int(px) >> shift, int(py) >> shift)  -- regda_7.py:3033 / :3195 `(preds / d).astype(int)`
int launch_decode(const float* heat, int n_maps, int H, int W, float* preds, float* maxvals, int32_t* idx,
                  int32_t* centres, int shift, cudaStream_t stream);

// ---- peer mailboxes (hp_peer.cu): slots[2 parities][world][kPeerWords int64] + 8 words (step counter) ----------
constexpr int kPeerWords = 64;     // int64 words per (parity, source) slot; word 63 is the flag
constexpr int kPeerMaxWorld = 16;
struct PeerLink {                  // passed by value to kernels that do the exchange themselves
    long long* mailbox[kPeerMaxWorld];  // base of every rank's mailbox as mapped in this process
    int rank, world;                    // world <= 1: no exchange
};
__device__ __forceinline__ long long* peer_slot(long long* base, int world, int parity, int src) {
    return base + (static_cast<size_t>(parity) * world + src) * kPeerWords;
}
__device__ __forceinline__ long long* peer_counter(long long* base, int world) {
    return base + static_cast<size_t>(2) * world * kPeerWords;
}

// the per-step exchange + finalise over peer mailboxes (hp_peer.cu); `overlap` != 0: programmatic dependent launch
int launch_finalize_peer(const long long* partial, void* const* mailboxes, int rank, int world, int K, long long seq,
                         long long* partial_out, double* result, int overlap, cudaStream_t stream);

}  // namespace hp
