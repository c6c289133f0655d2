// hp_peer.cu - the path's single collective without NCCL on the per-step path.
//
// The batch-sharded pipeline needs ONE all-reduce(sum) per step of a 52-element int64 vector (416 B;
// SURVEY.md 8e).  At a 28 us step the cost of that collective is pure latency and host launch overhead,
// so it is done by the finalise kernel itself over NVLink peer memory:
//   * every rank owns a MAILBOX (plain cudaMalloc memory, exported with CUDA IPC and mapped by all ranks
//     of the node): slots[2 parities][world][64 int64];
//   * hp_pipeline_finalize_peer on rank r writes its partial vector into slot[parity][r] of EVERY rank's
//     mailbox (P2P stores), fences at system scope, then writes the step's sequence number as the flag;
//   * it then waits (bounded spin, volatile system-scope loads) until all `world` flags of its own mailbox
//     show this sequence number, sums the `world` vectors in rank order and finalises exactly like
//     hp_pipeline_finalize.  Every entry is an integer, so all ranks get bit-identical results.
// The step number is either passed by the host or (seq = 0) counted on the device, which makes a captured CUDA
// graph of steps replayable.  Two parities suffice: a rank can be at most one step ahead of the slowest rank (it cannot finish step
// s+1 before it has received everybody's step-(s+1) vector).  Ranks run on different GPUs, so the kernels
// that wait for one another always execute concurrently; the spin is bounded (~2 s) and reports a timeout
// through the result vector instead of hanging.
#include <cstring>

#include "hp_common.cuh"
#include "hp_internal.cuh"
#include "hp_pipeline_common.cuh"

namespace hp {

struct PeerArgs {
    const long long* partial;           // this rank's partial vector [4+2K+6]
    long long* mailbox[kPeerMaxWorld];  // base of every rank's mailbox as mapped in this process
    int rank, world, K;
    long long seq;                      // 1, 2, 3, ... (one per step, identical on all ranks)
    long long* partial_out;             // nullable: the reduced vector
    double* result;                     // [4+K]
};

// 64 threads, a few registers, 1.2 KB of shared memory: always fits beside the resident pipeline blocks of a train
__global__ void __launch_bounds__(64) pipeline_finalize_peer_kernel(const PeerArgs a) {
    __shared__ long long s_total[4 + 2 * HP_MAX_K + 6];
    __shared__ int s_timeout;
    const int n = 4 + 2 * a.K + 6;
    // Programmatic dependent launch (no-ops without the launch attribute): the next pipeline launch of the train
    // may start right away; this kernel reads `partial` only once the pipeline launch before it has completed.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    // (the counter is read only now: the previous step's exchange kernel, which writes it, has completed)
    // step number: given by the host, or (seq == 0) counted on the device in the rank's own mailbox so that a
    // captured CUDA graph of steps can be replayed - all ranks run the same steps in the same order
    long long* counter = peer_counter(a.mailbox[a.rank], a.world);
    const long long seq = a.seq != 0 ? a.seq : *reinterpret_cast<volatile long long*>(counter) + 1;
    const int parity = static_cast<int>(seq & 1);
    if (threadIdx.x == 0) s_timeout = 0;
    // ---- send: my vector into slot[parity][rank] of every mailbox ------------------------------------
    for (int i = threadIdx.x; i < a.world * n; i += blockDim.x) {
        const int dst = i / n, w = i - dst * n;
        peer_slot(a.mailbox[dst], a.world, parity, a.rank)[w] = a.partial[w];
    }
    __syncthreads();
    if (threadIdx.x < a.world) {
        __threadfence_system();  // payload before flag, at system scope (peer GPUs)
        volatile long long* flag = peer_slot(a.mailbox[threadIdx.x], a.world, parity, a.rank) + (kPeerWords - 1);
        *flag = seq;
    }
    // ---- receive: wait for every source's flag in MY mailbox ---------------------------------------------
    if (threadIdx.x < a.world) {
        volatile long long* flag = peer_slot(a.mailbox[a.rank], a.world, parity, threadIdx.x) + (kPeerWords - 1);
        const long long t0 = clock64();
        while (*flag != seq) {
            if (clock64() - t0 > 4000000000ll) {  // ~2 s at 2 GHz: a peer never arrived
                s_timeout = 1;
                break;
            }
        }
        __threadfence_system();
    }
    __syncthreads();
    // ---- reduce in rank order ---------------------------------------------------------------------------------
    for (int w = threadIdx.x; w < n; w += blockDim.x) {
        long long t = 0;
        for (int src = 0; src < a.world; ++src)
            t += *reinterpret_cast<volatile long long*>(peer_slot(a.mailbox[a.rank], a.world, parity, src) + w);
        s_total[w] = t;
        if (a.partial_out) a.partial_out[w] = t;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        *counter = seq;  // every thread has read it (barriers above)
        pipeline_result_from_partial(s_total, a.K, a.result);
        if (s_timeout) a.result[0] = a.result[1] = __longlong_as_double(0x7ff8000000000000ll);
    }
}

}  // namespace hp

using namespace hp;

// ---- setup-time helpers (the only entry points of the library that own memory) -------------------------------
extern "C" HP_API size_t hp_peer_mailbox_bytes(int world) {
    return sizeof(long long) * (2 * static_cast<size_t>(world > 0 ? world : 1) * kPeerWords + 8);  // + step counter
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
                   4 + 2 * K + 6 < kPeerWords,
               HP_ERR_ARG, "hp_pipeline_finalize_peer: rank=%d world=%d K=%d seq=%lld", rank, world, K, seq);
    PeerArgs a{};
    a.partial = partial;
    for (int r = 0; r < world; ++r) {
        HP_REQUIRE(mailboxes[r], HP_ERR_NULL, "hp_pipeline_finalize_peer: mailbox %d is null", r);
        a.mailbox[r] = static_cast<long long*>(mailboxes[r]);
    }
    a.rank = rank; a.world = world; a.K = K; a.seq = seq;
    a.partial_out = partial_out; a.result = result;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(1);
    cfg.blockDim = dim3(64);
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
