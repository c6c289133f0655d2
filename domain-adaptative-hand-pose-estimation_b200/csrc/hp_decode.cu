// hp_decode.cu - argmax decode (a1) and PCK accuracy (a2).
//
// Replaces utils/keypoint_detection.py:7-35 (get_max_preds) and :38-92 (calc_dists, dist_acc,
// accuracy) of the reference, which run on the host after a full device->host copy.
//
// Kernel shape: one map per thread group; the group holds the whole map in registers
// (read from HBM once with 128-bit no-allocate loads), scans it in index order and merges
// (value, index) pairs with numpy's tie rules.  Roofline: HBM; algorithmic bytes per map =
// H*W*4 read (+12 written).
#include "hp_common.cuh"
#include "hp_decode.cuh"
#include "hp_dispatch.cuh"
#include "hp_internal.cuh"
#include "hp_decode_staged.cuh"

namespace hp {

template <int TPM, int NV, int MODE, int MPB>
__global__ void __launch_bounds__(TPM* MPB)
    decode_kernel(const float* __restrict__ heat, int n_maps, int HW, int W, float* __restrict__ preds,
                  float* __restrict__ maxvals, int32_t* __restrict__ idx_out, int32_t* __restrict__ centres,
                  int shift) {
    __shared__ Stats<0> scratch[TPM > 32 ? TPM / 32 + 1 : 1];
    // a kernel launched behind this one with the programmatic-serialisation attribute (the dense disparity kernel,
    // hp_regdisp_dense.cuh) may become resident as soon as every block of this grid has started; it reads this kernel's
    // output only after its own griddepcontrol.wait.  No effect on ordinary launches.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int g = threadIdx.x / TPM, t = threadIdx.x % TPM;
    const int map = blockIdx.x * MPB + g;
    if (map >= n_maps) return;  // TPM == 32 groups are whole warps; TPM > 32 implies MPB == 1
    const ArgMax a = decode_map<TPM, NV, MODE>(heat + static_cast<size_t>(map) * HW, HW, t, scratch);
    if (t == 0) {
        float px, py;
        decode_xy(a, W, px, py);
        if (preds) {
            preds[2 * map + 0] = px;
            preds[2 * map + 1] = py;
        }
        if (maxvals) maxvals[map] = a.v;
        if (idx_out) idx_out[map] = a.i;
        if (centres) {
            centres[2 * map + 0] = static_cast<int>(px) >> shift;
            centres[2 * map + 1] = static_cast<int>(py) >> shift;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// accuracy(output, target): decode both tensors + PCK counts + finalise by the last block
// ---------------------------------------------------------------------------------------------
template <int TPM, int NV, int MODE, int MPB>
__global__ void __launch_bounds__(TPM* MPB)
    accuracy_kernel(const float* __restrict__ output, const float* __restrict__ target, int n_maps, int K, int H, int W,
                    double thr, float* __restrict__ pred_xy, int32_t* __restrict__ counts_out,
                    double* __restrict__ acc_out, Workspace* __restrict__ ws) {
    __shared__ Stats<0> scratch[TPM > 32 ? TPM / 32 + 1 : 1];
    const int HW = H * W;
    const int g = threadIdx.x / TPM, t = threadIdx.x % TPM;
    const int map = blockIdx.x * MPB + g;
    if (map < n_maps) {
        const ArgMax ao = decode_map<TPM, NV, MODE>(output + static_cast<size_t>(map) * HW, HW, t, scratch);
        const ArgMax at = decode_map<TPM, NV, MODE>(target + static_cast<size_t>(map) * HW, HW, t, scratch);
        if (t == 0) {
            float px, py, tx, ty;
            decode_xy(ao, W, px, py);
            decode_xy(at, W, tx, ty);
            pred_xy[2 * map + 0] = px;
            pred_xy[2 * map + 1] = py;
            int valid, hit;
            pck_one(px, py, tx, ty, H, W, thr, valid, hit);
            const int k = map % K;
            if (valid) atomicAdd(&ws->counts[K + k], 1);
            if (hit) atomicAdd(&ws->counts[k], 1);
        }
    }
    if (last_block_arrives_writers(&ws->counter, gridDim.x, t == 0 && map < n_maps)) pck_publish(ws, K, counts_out, acc_out);
}

__global__ void pck_accumulate_kernel(const float* __restrict__ pred_xy, const float* __restrict__ tgt_xy, int n_maps,
                                      int K, int H, int W, double thr, int32_t* __restrict__ counts) {
    const int map = blockIdx.x * blockDim.x + threadIdx.x;
    if (map >= n_maps) return;
    int valid, hit;
    pck_one(pred_xy[2 * map], pred_xy[2 * map + 1], tgt_xy[2 * map], tgt_xy[2 * map + 1], H, W, thr, valid, hit);
    const int k = map % K;
    if (valid) atomicAdd(&counts[K + k], 1);
    if (hit) atomicAdd(&counts[k], 1);
}

__global__ void pck_finalize_kernel(const int32_t* __restrict__ counts, int K, double* __restrict__ acc_out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        int hits[HP_MAX_K], valid[HP_MAX_K];
        for (int k = 0; k < K; ++k) {
            hits[k] = counts[k];
            valid[k] = counts[K + k];
        }
        pck_finalize_serial(hits, valid, K, acc_out);
    }
}

struct DecodeLaunch {
    const float* heat;
    int n_maps, HW, W;
    float* preds;
    float* maxvals;
    int32_t* idx;
    int32_t* centres;
    int shift;
    cudaStream_t stream;
    template <int TPM, int NV, int MODE, int MPB>
    void run() const {
        const int grid = (n_maps + MPB - 1) / MPB;
        decode_kernel<TPM, NV, MODE, MPB>
            <<<grid, TPM * MPB, 0, stream>>>(heat, n_maps, HW, W, preds, maxvals, idx, centres, shift);
    }
};

struct AccuracyLaunch {
    const float* output;
    const float* target;
    int n_maps, K, H, W;
    double thr;
    float* pred_xy;
    int32_t* counts;
    double* acc;
    Workspace* ws;
    cudaStream_t stream;
    template <int TPM, int NV, int MODE, int MPB>
    void run() const {
        const int grid = (n_maps + MPB - 1) / MPB;
        accuracy_kernel<TPM, NV, MODE, MPB>
            <<<grid, TPM * MPB, 0, stream>>>(output, target, n_maps, K, H, W, thr, pred_xy, counts, acc, ws);
    }
};

int launch_decode(const float* heat, int n_maps, int H, int W, float* preds, float* maxvals, int32_t* idx,
                  int32_t* centres, int shift, cudaStream_t stream) {
    if (n_maps == 0) return HP_OK;
    DecodeLaunch l{heat, n_maps, H * W, W, preds, maxvals, idx, centres, shift, stream};
    dispatch_map_walk(H * W, aligned16(heat), l);
    return launch_status("decode");
}

}  // namespace hp

using namespace hp;

extern "C" HP_API int hp_argmax_decode(const float* heat, int n_maps, int H, int W, float* preds, float* maxvals,
                                int32_t* idx, hp_stream_t stream) {
    HP_REQUIRE(heat && preds && maxvals, HP_ERR_NULL, "hp_argmax_decode: null pointer");
    HP_REQUIRE(n_maps >= 0 && H > 0 && W > 0 && static_cast<long long>(H) * W < (1ll << 30), HP_ERR_SHAPE,
               "hp_argmax_decode: bad shape n_maps=%d H=%d W=%d", n_maps, H, W);
    HP_REQUIRE(aligned4(heat) && aligned4(preds) && aligned4(maxvals), HP_ERR_ALIGN, "hp_argmax_decode: misaligned");
    return launch_decode(heat, n_maps, H, W, preds, maxvals, idx, nullptr, 0, static_cast<cudaStream_t>(stream));
}

extern "C" HP_API int hp_pck_accumulate(const float* pred_xy, const float* tgt_xy, int B, int K, int H, int W, double thr,
                                 int32_t* counts, hp_stream_t stream) {
    HP_REQUIRE(pred_xy && tgt_xy && counts, HP_ERR_NULL, "hp_pck_accumulate: null pointer");
    HP_REQUIRE(B >= 0 && K > 0 && K <= HP_MAX_K && H > 0 && W > 0, HP_ERR_SHAPE, "hp_pck_accumulate: bad shape");
    const int n = B * K;
    if (n == 0) return HP_OK;
    pck_accumulate_kernel<<<(n + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(pred_xy, tgt_xy, n, K, H, W,
                                                                                          thr, counts);
    return launch_status("hp_pck_accumulate");
}

extern "C" HP_API int hp_pck_finalize(const int32_t* counts, int K, double* acc_out, hp_stream_t stream) {
    HP_REQUIRE(counts && acc_out, HP_ERR_NULL, "hp_pck_finalize: null pointer");
    HP_REQUIRE(K > 0 && K <= HP_MAX_K, HP_ERR_SHAPE, "hp_pck_finalize: K=%d out of range", K);
    HP_REQUIRE(aligned8(acc_out), HP_ERR_ALIGN, "hp_pck_finalize: acc_out must be 8-byte aligned");
    pck_finalize_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(counts, K, acc_out);
    return launch_status("hp_pck_finalize");
}

extern "C" HP_API int hp_accuracy(const float* output, const float* target, int B, int K, int H, int W, double thr,
                           float* pred_xy, int32_t* counts, double* acc_out, void* workspace, hp_stream_t stream) {
    HP_REQUIRE(output && target && pred_xy && counts && acc_out && workspace, HP_ERR_NULL, "hp_accuracy: null pointer");
    HP_REQUIRE(B > 0 && K > 0 && K <= HP_MAX_K && H > 0 && W > 0 && static_cast<long long>(H) * W < (1ll << 30),
               HP_ERR_SHAPE, "hp_accuracy: bad shape B=%d K=%d H=%d W=%d", B, K, H, W);
    HP_REQUIRE(aligned8(acc_out) && aligned8(workspace), HP_ERR_ALIGN, "hp_accuracy: misaligned output");
    {   // 64x64 maps: warp-private copy-engine stages (hp_decode_staged.cuh); everything else: block per map
        const int rc = launch_accuracy_staged(output, target, B * K, K, H, W, thr, pred_xy, counts, acc_out,
                                              static_cast<Workspace*>(workspace), static_cast<cudaStream_t>(stream));
        if (rc != 1) return rc;
    }
    AccuracyLaunch l{output, target, B * K, K, H, W, thr, pred_xy, counts, acc_out,
                     static_cast<Workspace*>(workspace), static_cast<cudaStream_t>(stream)};
    dispatch_map_walk(H * W, aligned16(output) && aligned16(target), l);
    return launch_status("hp_accuracy");
}
