// hp_decode.cuh - the argmax scan of one map by a TPM-thread group, over a pluggable tile source
// (global memory for get_max_preds; a fused multiscale map computed on the fly for config 4).
#pragma once
#include "hp_common.cuh"

namespace hp {

// Loader: void operator()(int tile, float fill, float4 (&v)[NV]) const - fills the registers of
// tile `tile` (layout of load_tile), out-of-range elements = fill.
template <int TPM, int NV, int MODE, class Loader>
__device__ __forceinline__ ArgMax decode_tiles(const Loader& load, int HW, int t, Stats<0>* scratch) {
    Stats<0> st;
    stats_init(st);
    const int ntiles = (MODE == WALK_EXACT) ? 1 : tiles_for<TPM, NV>(HW);
    float4 v[NV];
    for (int tile = 0; tile < ntiles; ++tile) {
        load(tile, -3.402823466e38f, v);
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            const int idx0 = tile * (TPM * NV * 4) + (j * TPM + t) * 4;
            am_scan4<false>(st.am, v[j], idx0);
            st.flag = fmaf(v[j].x, 0.f, st.flag);
            st.flag = fmaf(v[j].y, 0.f, st.flag);
            st.flag = fmaf(v[j].z, 0.f, st.flag);
            st.flag = fmaf(v[j].w, 0.f, st.flag);
        }
    }
    group_reduce<TPM, 0, true, false, false>(st, scratch);
    if (st.flag != st.flag) {
        // a NaN or an infinity is present (group-uniform): redo the scan with the exact numpy rules
        stats_init(st);
        for (int tile = 0; tile < ntiles; ++tile) {
            if (MODE != WALK_EXACT) load(tile, -INFINITY, v);
#pragma unroll
            for (int j = 0; j < NV; ++j) {
                const int idx0 = tile * (TPM * NV * 4) + (j * TPM + t) * 4;
                if (MODE == WALK_EXACT) {
                    am_scan4<true>(st.am, v[j], idx0);
                } else {
                    if (idx0 + 0 < HW) am_scan1<true>(st.am, v[j].x, idx0 + 0);
                    if (idx0 + 1 < HW) am_scan1<true>(st.am, v[j].y, idx0 + 1);
                    if (idx0 + 2 < HW) am_scan1<true>(st.am, v[j].z, idx0 + 2);
                    if (idx0 + 3 < HW) am_scan1<true>(st.am, v[j].w, idx0 + 3);
                }
            }
        }
        group_reduce<TPM, 0, true, false, false>(st, scratch);
    }
    return st.am;
}

template <int TPM, int NV, int MODE>
struct GlobalTileLoader {
    const float* map;
    int HW, t;
    __device__ __forceinline__ void operator()(int tile, float fill, float4 (&v)[NV]) const {
        load_tile<TPM, NV, MODE>(map, HW, tile, t, fill, v);
    }
};

template <int TPM, int NV, int MODE>
__device__ __forceinline__ ArgMax decode_map(const float* __restrict__ map, int HW, int t, Stats<0>* scratch) {
    GlobalTileLoader<TPM, NV, MODE> ld{map, HW, t};
    return decode_tiles<TPM, NV, MODE>(ld, HW, t, scratch);
}

// finalise PCK by the last block, ALL threads: one thread per counter (one L2 round trip instead of 2K
// dependent ones - the single-thread version cost ~15 us), then the ordered average from shared memory
__device__ __forceinline__ void pck_publish(Workspace* ws, int K, int32_t* counts_out, double* acc_out) {
    __shared__ int s_cnt[2 * HP_MAX_K];
    __shared__ double s_acc[HP_MAX_K];
    for (int i = threadIdx.x; i < 2 * K; i += blockDim.x) {
        const int v = *reinterpret_cast<volatile int*>(&ws->counts[i]);
        ws->counts[i] = 0;
        s_cnt[i] = v;
        if (counts_out) counts_out[i] = v;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        const int h = s_cnt[k], v = s_cnt[K + k];
        const double acc = v > 0 ? __ddiv_rn(static_cast<double>(h) * 1.0, static_cast<double>(v)) : -1.0;
        s_acc[k] = acc;
        acc_out[k] = acc;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double total = 0.0;
        int cnt = 0;
        for (int k = 0; k < K; ++k)
            if (s_acc[k] >= 0.0) {
                total = __dadd_rn(total, s_acc[k]);
                ++cnt;
            }
        acc_out[K] = cnt != 0 ? __ddiv_rn(total, static_cast<double>(cnt)) : 0.0;
        acc_out[K + 1] = static_cast<double>(cnt);
        ws->counter = 0;
    }
}

}  // namespace hp
