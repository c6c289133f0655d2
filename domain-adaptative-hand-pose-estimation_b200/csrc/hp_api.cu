// hp_api.cu - library-level entry points and the error plumbing of libhp_b200.so.
#include <cstdarg>
#include <cstdio>

#include "hp_common.cuh"

namespace hp {

static thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int launch_status(const char* what) {
    const cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) return HP_OK;
    snprintf(g_err, sizeof(g_err), "%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
    return static_cast<int>(e);
}

}  // namespace hp

extern "C" HP_API int hp_version(void) { return 100; /* 0.1.0 */ }

extern "C" HP_API const char* hp_last_error(void) { return hp::g_err; }

extern "C" HP_API int hp_device_sm_count(void) {
    int dev = 0, sms = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) {
        hp::fail(-1, "hp_device_sm_count: %s", cudaGetErrorString(e));
        return -static_cast<int>(e);
    }
    return sms;
}

extern "C" HP_API size_t hp_workspace_bytes(int n_maps, int K) {
    (void)K;
    // Workspace header (block counter, PCK counters, 64-bit loss accumulators) + slack; kernels that need
    // per-map scratch keep it in shared memory
    (void)n_maps;
    // + at byte 1024 the record of a deferred cross-GPU exchange (hp_internal.cuh) and at byte 2048 the self-certifying
    // accumulators of the fused pipeline kernel (hp_pipeline_bulk.cuh: 4 copies of 160 words); all zero between launches
    return 2048 + 4 * 160 * sizeof(unsigned long long);
}
