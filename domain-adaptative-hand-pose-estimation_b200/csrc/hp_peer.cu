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

namespace hp {

struct PeerArgs {
    const long long* partial;           // this rank's partial vector [4+2K+6]
    PeerLink link;
    int K;
    long long* partial_out;             // nullable: the reduced vector
    double* result;                     // [4+K]
};

// one warp, a few registers, 9 KB of shared memory: always fits beside the resident pipeline blocks of a train
__global__ void __launch_bounds__(32) pipeline_finalize_peer_kernel(const PeerArgs a) {
    __shared__ long long s_total[4 + 2 * HP_MAX_K + 6];
    __shared__ long long s_scratch[kPeerScratchWords];
    const int n = 4 + 2 * a.K + 6, lane = threadIdx.x;
    // Programmatic dependent launch (no-ops without the launch attribute): the next pipeline launch of the train
    // may start right away; this kernel reads `partial` only once the pipeline launch before it has completed.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    for (int w = lane; w < n; w += 32) s_total[w] = a.partial[w];
    __syncwarp();
    const int timeout = peer_exchange_warp(a.link, s_total, s_scratch, n, lane);
    if (a.partial_out)
        for (int w = lane; w < n; w += 32) a.partial_out[w] = s_total[w];
    if (lane == 0) {
        pipeline_result_from_partial(s_total, a.K, a.result);
        if (timeout) a.result[0] = a.result[1] = __longlong_as_double(0x7ff8000000000000ll);
    }
}

}  // namespace hp

using namespace hp;

// ---- setup-time helpers (the only entry points of the library that own memory) -------------------------------
extern "C" HP_API size_t hp_peer_mailbox_bytes(int world) {
    return sizeof(unsigned long long) * (2 * static_cast<size_t>(world > 0 ? world : 1) * kPeerSlotEntries + 8);  // + step counter
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
int launch_finalize_peer(const long long* partial, void* const* mailboxes, int rank, int world, int K, long long seq,
                         long long* partial_out, double* result, int overlap, cudaStream_t stream) {
    HP_REQUIRE(partial && mailboxes && result, HP_ERR_NULL, "hp_pipeline_finalize_peer: null pointer");
    HP_REQUIRE(world > 0 && world <= kPeerMaxWorld && rank >= 0 && rank < world && K > 0 && K <= HP_MAX_K && seq >= 0 &&
                   peer_shape_ok(K, world),
               HP_ERR_ARG, "hp_pipeline_finalize_peer: rank=%d world=%d K=%d seq=%lld", rank, world, K, seq);
    HP_REQUIRE(seq == 0, HP_ERR_ARG, "hp_pipeline_finalize_peer: seq=%lld (the step is counted on the device: pass 0)", seq);
    PeerArgs a{};
    a.partial = partial;
    for (int r = 0; r < world; ++r) {
        HP_REQUIRE(mailboxes[r], HP_ERR_NULL, "hp_pipeline_finalize_peer: mailbox %d is null", r);
        a.link.mailbox[r] = static_cast<unsigned long long*>(mailboxes[r]);
    }
    a.link.rank = rank; a.link.world = world; a.K = K;
    a.partial_out = partial_out; a.result = result;
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
    const cudaError_t e = cudaLaunchKernelEx(&cfg, pipeline_finalize_peer_kernel, a);
    if (e != cudaSuccess) return fail(static_cast<int>(e), "hp_pipeline_finalize_peer: %s", cudaGetErrorString(e));
    return launch_status("hp_pipeline_finalize_peer");
}
}  // namespace hp

extern "C" HP_API int hp_pipeline_finalize_peer(const int64_t* partial, void* const* mailboxes, int rank, int world,
                                                int K, int64_t seq, int64_t* partial_out, double* result,
                                                hp_stream_t stream) {
    return launch_finalize_peer(reinterpret_cast<const long long*>(partial), mailboxes, rank, world, K,
                                static_cast<long long>(seq), reinterpret_cast<long long*>(partial_out), result, 0,
                                static_cast<cudaStream_t>(stream));
}
