// hp_peer.cu - the path's single collective without NCCL on the per-step path.
//
// The batch-sharded pipeline needs ONE all-reduce(sum) per step of a 52-element int64 vector (416 B;
// SURVEY.md 8e).  At a 12 us step the cost of that collective is pure latency and host launch overhead,
// so it is done over NVLink peer memory by one warp - normally the last block of the fused pipeline kernel
// itself (hp_pipeline_bulk.cuh), or the small kernel below for shapes that kernel does not serve:
//   * every rank owns a MAILBOX (plain cudaMalloc memory, exported with CUDA IPC and mapped by all ranks
//     of the node);
//   * the warp stores its vector into every rank's mailbox as tagged 8-byte entries (32 bits of payload +
//     the 32-bit step number: no fence, no separate flag), polls its own mailbox until every entry of every
//     source carries the current step, sums the vectors in rank order and finalises exactly like
//     hp_pipeline_finalize.  Every entry is an integer, so all ranks get bit-identical results.
// Layout, protocol and the two-parity argument: hp_internal.cuh.  Ranks run on different GPUs, so the kernels
// that wait for one another always execute concurrently; the wait is bounded (~2 s) and reports a timeout
// through the result vector instead of hanging.
#include <cstring>

#include "hp_common.cuh"
#include "hp_internal.cuh"
#include "hp_pipeline_common.cuh"
#include "hp_peer_step.cuh"

namespace hp {

struct PeerArgs {
    const long long* partial;           // this rank's partial vector [4+2K+6]
    PeerLink link;
    int K;
    long long* partial_out;             // nullable: the reduced vector
    double* result;                     // [4+K]
    Workspace* ws;                      // where a deferred step is remembered (a private dummy when the caller gave none)
    int defer;
    // hp_pck_finalize_peer: integer PCK counts instead of a partial vector
    const int* counts;                  // [2K] hits, valid of this rank
    int* counts_out;                    // [2K] totals
    double* acc_out;                    // [K+2] acc[K], avg_acc, cnt
};

// one warp, a few registers, ~11 KB of shared memory: always fits beside the resident pipeline blocks of a train
__global__ void __launch_bounds__(32) pipeline_finalize_peer_kernel(const PeerArgs a) {
    __shared__ long long s_total[4 + 2 * HP_MAX_K + 6];
    __shared__ long long s_scratch[kPeerScratchWords];
    __shared__ double s_acc[HP_MAX_K];
    __shared__ Workspace s_dummy_ws[4];  // >= kPendOffsetBytes + sizeof(PeerPending): a pending record that is never valid
    const int n = 4 + 2 * a.K + 6, lane = threadIdx.x;
    // Programmatic dependent launch (no-ops without the launch attribute): the next pipeline launch of the train
    // may start right away; this kernel reads `partial` only once the pipeline launch before it has completed.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    Workspace* ws = a.ws;
    if (!ws) {
        for (int i = lane; i < static_cast<int>(sizeof(s_dummy_ws) / 8); i += 32) reinterpret_cast<long long*>(s_dummy_ws)[i] = 0;
        __syncwarp();
        ws = s_dummy_ws;
    }
    for (int w = lane; w < n; w += 32) s_total[w] = a.partial[w];
    __syncwarp();
    peer_step_warp(a.link, ws, s_total, s_scratch, s_acc, a.K, a.partial_out, a.result, a.ws ? a.defer : 0, lane);
}

// the last step of a deferred train: collect and finalise what is still outstanding
__global__ void __launch_bounds__(32) pipeline_flush_peer_kernel(const PeerArgs a) {
    __shared__ long long s_scratch[kPeerScratchWords];
    __shared__ double s_acc[HP_MAX_K];
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    peer_flush_warp(a.link, a.ws, s_scratch, s_acc, threadIdx.x);
}

// configs[3]: sum of the 2K integer PCK counts over the ranks + the accuracy closure (keypoint_detection.py:80-90)
__global__ void __launch_bounds__(32) pck_finalize_peer_kernel(const PeerArgs a) {
    __shared__ long long s_total[4 + 2 * HP_MAX_K + 6];
    __shared__ long long s_scratch[kPeerScratchWords];
    __shared__ double s_acc[HP_MAX_K];
    __shared__ double s_result[4 + HP_MAX_K];
    __shared__ Workspace s_dummy_ws[4];
    const int K = a.K, n = 4 + 2 * K + 6, lane = threadIdx.x;
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    for (int i = lane; i < static_cast<int>(sizeof(s_dummy_ws) / 8); i += 32) reinterpret_cast<long long*>(s_dummy_ws)[i] = 0;
    for (int w = lane; w < n; w += 32) s_total[w] = (w >= 4 && w < 4 + 2 * K) ? a.counts[w - 4] : 0;
    __syncwarp();
    peer_step_warp(a.link, s_dummy_ws, s_total, s_scratch, s_acc, K, nullptr, s_result, 0, lane);
    const bool bad = s_result[3] != s_result[3];  // poisoned by a timeout
    for (int i = lane; i < 2 * K; i += 32) a.counts_out[i] = bad ? -1 : static_cast<int>(s_total[4 + i]);
    for (int k = lane; k < K; k += 32) a.acc_out[k] = s_result[4 + k];
    if (lane == 0) {
        a.acc_out[K] = s_result[2];
        a.acc_out[K + 1] = s_result[3];
    }
}

}  // namespace hp

using namespace hp;

// ---- setup-time helpers (the only entry points of the library that own memory) -------------------------------
extern "C" HP_API size_t hp_peer_mailbox_bytes(int world) {
    return sizeof(unsigned long long) * (static_cast<size_t>(kPeerRing) * (world > 0 ? world : 1) * kPeerSlotEntries + 8);  // + step counter
}

extern "C" HP_API int hp_peer_alloc(int world, void** mailbox) {
    HP_REQUIRE(mailbox && world > 0 && world <= kPeerMaxWorld, HP_ERR_ARG, "hp_peer_alloc: world=%d", world);
    const size_t bytes = hp_peer_mailbox_bytes(world);
    cudaError_t e = cudaMalloc(mailbox, bytes);
    if (e == cudaSuccess) e = cudaMemset(*mailbox, 0, bytes);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) return fail(static_cast<int>(e), "hp_peer_alloc: %s", cudaGetErrorString(e));
    return HP_OK;
}

extern "C" HP_API int hp_peer_free(void* mailbox) {
    const cudaError_t e = cudaFree(mailbox);
    if (e != cudaSuccess) return fail(static_cast<int>(e), "hp_peer_free: %s", cudaGetErrorString(e));
    return HP_OK;
}

/* handle: 64 bytes (cudaIpcMemHandle_t) */
extern "C" HP_API int hp_peer_export(void* mailbox, void* handle64) {
    HP_REQUIRE(mailbox && handle64, HP_ERR_NULL, "hp_peer_export: null pointer");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    const cudaError_t e = cudaIpcGetMemHandle(static_cast<cudaIpcMemHandle_t*>(handle64), mailbox);
    if (e != cudaSuccess) return fail(static_cast<int>(e), "hp_peer_export: %s", cudaGetErrorString(e));
    return HP_OK;
}

extern "C" HP_API int hp_peer_import(const void* handle64, void** mapped) {
    HP_REQUIRE(handle64 && mapped, HP_ERR_NULL, "hp_peer_import: null pointer");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    const cudaError_t e = cudaIpcOpenMemHandle(mapped, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) return fail(static_cast<int>(e), "hp_peer_import: %s", cudaGetErrorString(e));
    return HP_OK;
}

extern "C" HP_API int hp_peer_close(void* mapped) {
    const cudaError_t e = cudaIpcCloseMemHandle(mapped);
    if (e != cudaSuccess) return fail(static_cast<int>(e), "hp_peer_close: %s", cudaGetErrorString(e));
    return HP_OK;
}

// ---- per step -----------------------------------------------------------------------------------------------------
namespace hp {
static int fill_link(const char* who, PeerLink& link, void* const* mailboxes, int rank, int world, int K) {
    HP_REQUIRE(mailboxes, HP_ERR_NULL, "%s: null mailbox table", who);
    HP_REQUIRE(world > 0 && world <= kPeerMaxWorld && rank >= 0 && rank < world && K > 0 && K <= HP_MAX_K &&
                   peer_shape_ok(K, world),
               HP_ERR_ARG, "%s: rank=%d world=%d K=%d (K <= 27 when sharded)", who, rank, world, K);
    for (int r = 0; r < world; ++r) {
        HP_REQUIRE(mailboxes[r], HP_ERR_NULL, "%s: mailbox %d is null", who, r);
        link.mailbox[r] = static_cast<unsigned long long*>(mailboxes[r]);
    }
    link.rank = rank;
    link.world = world;
    return HP_OK;
}
template <typename Kernel>
static int launch_one_warp(const char* who, Kernel kernel, const PeerArgs& a, int overlap, cudaStream_t stream) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(1);
    cfg.blockDim = dim3(32);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = overlap ? 1 : 0;
    const cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, a);
    if (e != cudaSuccess) return fail(static_cast<int>(e), "%s: %s", who, cudaGetErrorString(e));
    return launch_status(who);
}

int launch_finalize_peer(const long long* partial, void* const* mailboxes, int rank, int world, int K, long long seq,
                         long long* partial_out, double* result, int overlap, void* workspace, int defer,
                         cudaStream_t stream) {
    HP_REQUIRE(partial && result, HP_ERR_NULL, "hp_pipeline_finalize_peer: null pointer");
    HP_REQUIRE(seq == 0, HP_ERR_ARG, "hp_pipeline_finalize_peer: seq=%lld (the step is counted on the device: pass 0)", seq);
    HP_REQUIRE(!defer || workspace, HP_ERR_NULL, "hp_pipeline_finalize_peer: a deferred step needs the workspace");
    PeerArgs a{};
    if (int rc = fill_link("hp_pipeline_finalize_peer", a.link, mailboxes, rank, world, K)) return rc;
    a.partial = partial; a.K = K; a.partial_out = partial_out; a.result = result;
    a.ws = static_cast<Workspace*>(workspace); a.defer = defer;
    return launch_one_warp("hp_pipeline_finalize_peer", pipeline_finalize_peer_kernel, a, overlap, stream);
}
}  // namespace hp

extern "C" HP_API int hp_pipeline_finalize_peer(const int64_t* partial, void* const* mailboxes, int rank, int world,
                                                int K, int64_t seq, int64_t* partial_out, double* result,
                                                hp_stream_t stream) {
    return launch_finalize_peer(reinterpret_cast<const long long*>(partial), mailboxes, rank, world, K,
                                static_cast<long long>(seq), reinterpret_cast<long long*>(partial_out), result, 0, nullptr, 0,
                                static_cast<cudaStream_t>(stream));
}

extern "C" HP_API int hp_pipeline_flush_peer(void* workspace, void* const* mailboxes, int rank, int world,
                                             hp_stream_t stream) {
    HP_REQUIRE(workspace, HP_ERR_NULL, "hp_pipeline_flush_peer: null workspace");
    PeerArgs a{};
    if (int rc = fill_link("hp_pipeline_flush_peer", a.link, mailboxes, rank, world, 1)) return rc;
    a.ws = static_cast<Workspace*>(workspace);
    return launch_one_warp("hp_pipeline_flush_peer", pipeline_flush_peer_kernel, a, 1, static_cast<cudaStream_t>(stream));
}

extern "C" HP_API int hp_pck_finalize_peer(const int32_t* counts, void* const* mailboxes, int rank, int world, int K,
                                           int32_t* counts_out, double* acc_out, hp_stream_t stream) {
    HP_REQUIRE(counts && counts_out && acc_out, HP_ERR_NULL, "hp_pck_finalize_peer: null pointer");
    PeerArgs a{};
    if (int rc = fill_link("hp_pck_finalize_peer", a.link, mailboxes, rank, world, K)) return rc;
    a.K = K; a.counts = counts; a.counts_out = counts_out; a.acc_out = acc_out;
    return launch_one_warp("hp_pck_finalize_peer", pck_finalize_peer_kernel, a, 1, static_cast<cudaStream_t>(stream));
}
