// hp_peer_step.cuh - one step of the path's single collective, executed by ONE warp: the publisher of the fused
// pipeline kernel (hp_pipeline_bulk.cuh) or the small stand-alone kernels of hp_peer.cu.  Protocol: hp_internal.cuh.
#pragma once
#include "hp_internal.cuh"
#include "hp_pipeline_common.cuh"

namespace hp {

// result = { mse, kl, avg_acc, cnt, acc[K] } from a partial vector in shared memory, by one warp
// (acc[k] in parallel, the ordered average over the joints by lane 0 - keypoint_detection.py:80-90)
__device__ __forceinline__ void warp_result_from_partial(const long long* vec, int K, double* result, double* acc_scratch,
                                                         int lane) {
    for (int k = lane; k < K; k += 32) {
        const long long h = vec[4 + k], v = vec[4 + K + k];
        const double acc = v > 0 ? __ddiv_rn(static_cast<double>(h) * 1.0, static_cast<double>(v)) : -1.0;
        acc_scratch[k] = acc;
        result[4 + k] = acc;
    }
    __syncwarp();
    if (lane == 0) {
        double total = 0.0;
        int cnt = 0;
        for (int k = 0; k < K; ++k)
            if (acc_scratch[k] >= 0.0) {
                total = __dadd_rn(total, acc_scratch[k]);
                ++cnt;
            }
        const long long* cls = vec + 4 + 2 * K;
        result[0] = loss_from_fx(vec[0], cls[0], cls[1], cls[2], vec[2]);
        result[1] = loss_from_fx(vec[1], cls[3], cls[4], cls[5], vec[2]);
        result[2] = cnt != 0 ? __ddiv_rn(total, static_cast<double>(cnt)) : 0.0;
        result[3] = static_cast<double>(cnt);
    }
    __syncwarp();
}
// a peer never arrived: the totals are incomplete - poison EVERY published number (losses, acc[], avg_acc, cnt) and flag
// the vector (every entry -1, so n_maps == -1) so that no caller can mistake them for a result
__device__ __forceinline__ void warp_poison(long long* partial, double* result, int n, int K, int lane) {
    if (result)
        for (int i = lane; i < 4 + K; i += 32) result[i] = __longlong_as_double(0x7ff8000000000000ll);
    if (partial)
        for (int i = lane; i < n; i += 32) partial[i] = -1;
}

// ONE warp.  vec[0..n): this rank's partial vector of the current step, in shared memory (clobbered).
//   defer == 0: exchange now; totals -> partial_out / result of THIS step.
//   defer != 0: send only; then collect and finalise the PENDING step (if any) into the buffers it recorded, and record
//               this step's buffers as pending (hp_internal.cuh "deferred exchange").
// A pending step is always completed first, so synchronous and deferred steps may be mixed freely.
// `scratch`: kPeerScratchWords int64, `acc_scratch`: HP_MAX_K doubles, both shared memory.
__device__ __forceinline__ void peer_step_warp(const PeerLink& link, Workspace* ws, long long* vec, long long* scratch,
                                               double* acc_scratch, int K, long long* partial_out, double* result,
                                               int defer, int lane) {
    const int n = 4 + 2 * K + 6;
    unsigned long long* counter = peer_counter(link.mailbox[link.rank], link.world);
    PeerPending* pend = peer_pending(ws);
    const unsigned long long done = peer_load(counter);  // steps fully sent by this rank so far
    const unsigned long long seq = done + 1ull;
    PeerPending old;
    old.partial = pend->partial; old.result = pend->result; old.seq = pend->seq; old.K = pend->K; old.valid = pend->valid;
    peer_send(link, vec, n, seq, lane);  // first: the stores travel while the pending step is collected
    int timeout = 0;
    if (old.valid) {
        const int on = 4 + 2 * old.K + 6;
        long long* tot = scratch + kPeerRecvWords;  // the vector-sized tail of the scratch area
        timeout = peer_recv_sum(link, tot, scratch, on, old.seq, lane);
        if (old.partial)
            for (int i = lane; i < on; i += 32) old.partial[i] = tot[i];
        if (old.result) warp_result_from_partial(tot, old.K, old.result, acc_scratch, lane);
        if (timeout) warp_poison(old.partial, old.result, on, old.K, lane);
    }
    if (defer) {
        if (lane == 0) {
            pend->partial = partial_out; pend->result = result; pend->seq = seq; pend->K = K; pend->valid = 1;
        }
    } else {
        const int t2 = peer_recv_sum(link, vec, scratch, n, seq, lane);
        if (partial_out)
            for (int i = lane; i < n; i += 32) partial_out[i] = vec[i];
        if (result) warp_result_from_partial(vec, K, result, acc_scratch, lane);
        if (t2) warp_poison(partial_out, result, n, K, lane);
        timeout |= t2;
        if (lane == 0 && old.valid) pend->valid = 0;
    }
    // a timed-out step is NOT counted: the next step re-uses its number, so ranks that did complete it and ranks that
    // did not cannot drift apart silently (the caller sees the poisoned result and stops)
    if (lane == 0 && !timeout) peer_store(counter, seq);
    __syncwarp();
}

// ---- the same step split over two warps of one block (the publisher block of the fused pipeline kernel) ---------------
// The pending step's collection (read the record, poll this rank's mailbox for the vectors every rank sent a step ago,
// finalise into the remembered buffers: two dependent L2 round trips + the writes) does not depend on THIS step's totals,
// so a second warp does it while the first one is still collecting the blocks' sums; the two meet at one named barrier
// before the record is overwritten.  At 8 GPUs this takes ~1.5 us off every step of a train.
//   warp B (any time after the previous grid has completed): peer_pending_warp(...)   then  named barrier
//   warp A (when this rank's vector is in `vec`):            named barrier            then  peer_send_warp(...)
__device__ __forceinline__ void peer_pending_warp(const PeerLink& link, Workspace* ws, long long* scratch, double* acc_scratch,
                                                  int lane) {
    PeerPending* pend = peer_pending(ws);
    PeerPending old;
    old.partial = pend->partial; old.result = pend->result; old.seq = pend->seq; old.K = pend->K; old.valid = pend->valid;
    if (!old.valid) return;
    const int on = 4 + 2 * old.K + 6;
    long long* tot = scratch + kPeerRecvWords;
    const int timeout = peer_recv_sum<16>(link, tot, scratch, on, old.seq, lane);  // <= 16 pairs per lane: ONE round of loads
    if (old.partial)
        for (int i = lane; i < on; i += 32) old.partial[i] = tot[i];
    if (old.result) warp_result_from_partial(tot, old.K, old.result, acc_scratch, lane);
    if (timeout) warp_poison(old.partial, old.result, on, old.K, lane);
    __syncwarp();
}
// `done`: this rank's step counter, loaded by the caller any time after the previous grid has completed
__device__ __forceinline__ void peer_send_warp(const PeerLink& link, Workspace* ws, long long* vec, long long* scratch,
                                               double* acc_scratch, int K, long long* partial_out, double* result,
                                               int defer, unsigned long long done, int lane) {
    const int n = 4 + 2 * K + 6;
    unsigned long long* counter = peer_counter(link.mailbox[link.rank], link.world);
    PeerPending* pend = peer_pending(ws);
    const unsigned long long seq = done + 1ull;
    peer_send(link, vec, n, seq, lane);
    int timeout = 0;
    if (defer) {
        if (lane == 0) {
            pend->partial = partial_out; pend->result = result; pend->seq = seq; pend->K = K; pend->valid = 1;
        }
    } else {
        timeout = peer_recv_sum(link, vec, scratch, n, seq, lane);
        if (partial_out)
            for (int i = lane; i < n; i += 32) partial_out[i] = vec[i];
        if (result) warp_result_from_partial(vec, K, result, acc_scratch, lane);
        if (timeout) warp_poison(partial_out, result, n, K, lane);
        if (lane == 0) pend->valid = 0;
    }
    if (lane == 0 && !timeout) peer_store(counter, seq);
    __syncwarp();
}

// ONE warp: complete the pending step of a deferred train (its vectors were sent by the step itself).
__device__ __forceinline__ void peer_flush_warp(const PeerLink& link, Workspace* ws, long long* scratch, double* acc_scratch,
                                                int lane) {
    PeerPending* pend = peer_pending(ws);
    if (!pend->valid) return;
    const unsigned long long done = pend->seq;
    const int K = pend->K, n = 4 + 2 * K + 6;
    long long* partial = pend->partial;
    double* result = pend->result;
    long long* tot = scratch + kPeerRecvWords;
    const int timeout = peer_recv_sum(link, tot, scratch, n, done, lane);
    if (partial)
        for (int i = lane; i < n; i += 32) partial[i] = tot[i];
    if (result) warp_result_from_partial(tot, K, result, acc_scratch, lane);
    if (timeout) warp_poison(partial, result, n, K, lane);
    __syncwarp();
    if (lane == 0) pend->valid = 0;
}

}  // namespace hp
