/*
 * hp_b200.h - C ABI of libhp_b200.so: the B200 (sm_100a) heatmap keypoint hot path.
 *
 * This is the drop-in boundary.  The reference has no FFI - its
 * "operator API" is a set of Python callables - so each entry point below names the
 * reference function (file:line, relative to the reference tree) whose arithmetic it
 * replaces; the Python layer in domain-adaptative-hand-pose-estimation_b200/ binds these
 * through ctypes and re-exposes the reference's own signatures.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name starts with h_ (pinned host memory);
 *   - heatmaps are contiguous NCHW float32: map (b,k) starts at ((b*K+k)*H*W);
 *   - the caller owns all memory, including workspaces; nothing is retained after a call;
 *   - every launch is asynchronous on the caller's stream (hp_stream_t == cudaStream_t);
 *     only the *_host entry points synchronise (they return host-visible results);
 *   - return 0 on success, <0 for an argument error (HP_ERR_*), >0 for a cudaError_t;
 *     hp_last_error() returns a thread-local description; no C++ exception crosses;
 *   - a workspace must be zero-filled once before its first use; calls leave it zeroed.
 *   - Gaussian values come from a caller-supplied table tab[d2] = exp(-d2 / (2 sigma^2)),
 *     d2 = 0 .. 2*tmp*tmp (float32, device memory).  The host builds it with the reference's
 *     own numpy expression (uda/dataset/util.py:49-54) so generated maps are bit-equal to the
 *     reference on the same host; `tmp` is the integer half-width of the pasted patch.
 */
#ifndef HP_B200_H
#define HP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* hp_stream_t; /* cudaStream_t */

#if defined(__GNUC__)
#define HP_API __attribute__((visibility("default")))
#else
#define HP_API
#endif

#define HP_OK             0
#define HP_ERR_NULL      -1 /* required pointer is NULL */
#define HP_ERR_SHAPE     -2 /* non-positive or unsupported extent */
#define HP_ERR_ALIGN     -3 /* pointer not aligned for its element type */
#define HP_ERR_ARG       -4 /* bad enum / scalar argument */

#define HP_MAX_K        64  /* joints per sample supported by the per-sample kernels */

/* loss selection bits for hp_pipeline_fused */
#define HP_LOSS_MSE      1
#define HP_LOSS_KL       2
#define HP_PARTIAL_LEN(K) (4 + 2 * (K) + 6) /* int64 elements of the pipeline partial vector */
#define HP_LOSS_FX_SHIFT 40 /* fixed-point scale of the loss sums: value * 2^40 */

/* pseudo-label variants (centre = decoded coordinate >> shift, on an oh x ow map) */
#define HP_PLG_BASE      0 /* PseudoLabelGenerator   uda/model/regda_4.py:17-86 : gf = clip(sum_{j!=k} gt_j) */
#define HP_PLG_ONE_MINUS 1 /* PseudoLabelGenerator01/03 uda/model/regda_7.py:2956-3201 : gf = clip(1-10 gt) */

/* regression-disparity variants */
#define HP_RD_BASE       0 /* RegressionDisparity    uda/model/regda_4.py:89-143  */
#define HP_RD_X1         1 /* RegressionDisparityx1  uda/model/regda_7.py:3206-3268 */
#define HP_RD_X5         2 /* RegressionDisparityx5  uda/model/regda_7.py:3485-3561 */
#define HP_RD_X6         3 /* RegressionDisparityx6  uda/model/regda_7.py:3564-3632 */
#define HP_RD_RD4        4 /* RegressionDisparity4   uda/model/regda_4.py:299-356: clip(clip(sum gt) - 10 gt), no fused map, no normalisation */
#define HP_MODE_MIN      0
#define HP_MODE_MAX      1

/* gradient layouts accepted by the backward kernels */
#define HP_GRAD_SCALAR     0 /* upstream gradient of the 'mean' scalar                  */
#define HP_GRAD_PER_MAP    1 /* upstream gradient [B*K] (MSE 'none')                     */
#define HP_GRAD_PER_SAMPLE 2 /* upstream gradient [B]   (KL  'none' = mean over joints)  */

/* ---- library ------------------------------------------------------------------------ */
int         hp_version(void);            /* 10000*major + 100*minor + patch */
const char* hp_last_error(void);
int         hp_device_sm_count(void);    /* SMs of the current device (148 on B200), <0 on error */

/* bytes of zero-initialised workspace every op below accepts (one size fits all ops) */
size_t      hp_workspace_bytes(int n_maps, int K);

/* ---- a1: decode.  utils/keypoint_detection.py:7-35 (get_max_preds) ------------------- */
/* preds [n_maps,2] = (x,y) zeroed where max <= 0 ; maxvals [n_maps] ; first index wins ties,
 * NaN counts as the maximum.  idx (nullable) receives the flat int32 argmax. */
int hp_argmax_decode(const float* heat, int n_maps, int H, int W,
                     float* preds, float* maxvals, int32_t* idx, hp_stream_t stream);

/* ---- f4: soft-argmax.  utils/keypoint_detection.py:209-239 (compute_uv_from_heatmaps3) ---- */
/* out_uv [n_maps,2] = scale * (E[col], E[row]) under softmax(beta * heat) over H*W
 * (the reference: beta = 100, scale = 4). */
int hp_soft_argmax(const float* heat, int n_maps, int H, int W, float beta, float scale,
                   float* out_uv, hp_stream_t stream);

/* ---- a2: PCK.  utils/keypoint_detection.py:38-92 (calc_dists, dist_acc, accuracy) ---- */
/* From decoded coordinates: counts[0..K) += hits, counts[K..2K) += valid (int32). */
int hp_pck_accumulate(const float* pred_xy, const float* tgt_xy, int B, int K, int H, int W,
                      double thr, int32_t* counts, hp_stream_t stream);
/* acc_out[0..K) = hits/valid or -1 ; acc_out[K] = avg_acc ; acc_out[K+1] = cnt  (float64). */
int hp_pck_finalize(const int32_t* counts, int K, double* acc_out, hp_stream_t stream);
/* accuracy(output, target): decode both tensors, count, finalise - one launch.
 * pred_xy [B*K,2] float32 ; counts [2K] int32 (overwritten) ; acc_out [K+2] float64. */
int hp_accuracy(const float* output, const float* target, int B, int K, int H, int W, double thr,
                float* pred_xy, int32_t* counts, double* acc_out, void* workspace,
                hp_stream_t stream);

/* ---- a5: target generation.  uda/dataset/util.py:9-68 (generate_target), batched ------ */
/* joints float64 [B*K,2] image px ; vis float32 [B*K] ; stride = image_size / heatmap_size
 * (float64, computed by the host exactly as util.py:36) ; target [B*K,H,W] ; weight [B*K]. */
int hp_gaussian_target(const double* joints, const float* vis, int n_maps, int H, int W,
                       double stride_x, double stride_y, int tmp, const float* tab,
                       float* target, float* weight, hp_stream_t stream);

/* ---- a3: JointsMSELoss.  uda/model/loss.py:27-65 -------------------------------------- */
/* per_map [B*K] = mean_hw(0.5 w (p-t)^2) ('none') ; mean (nullable) = mean over all elements. */
int hp_mse_fwd(const float* output, const float* target, const float* weight /*nullable [B*K]*/,
               int B, int K, int HW, float* per_map, float* mean, void* workspace,
               hp_stream_t stream);
int hp_mse_bwd(const float* output, const float* target, const float* weight,
               const float* grad_out, int grad_kind, int B, int K, int HW, float* grad_in,
               hp_stream_t stream);

/* ---- a4: JointsKLLoss.  uda/model/loss.py:115-158 ------------------------------------- */
/* per_map [B*K] = w * KL(q || softmax(p)) ; per_sample (nullable) [B] = mean over K ('none');
 * mean (nullable) scalar ; stats [B*K,2] = (logsumexp, sum(t+eps)) saved for the backward. */
int hp_kl_fwd(const float* output, const float* target, const float* weight, float epsilon,
              int B, int K, int HW, float* per_map, float* per_sample, float* mean, float* stats,
              void* workspace, hp_stream_t stream);
int hp_kl_bwd(const float* output, const float* target, const float* weight, float epsilon,
              const float* stats, const float* grad_out, int grad_kind, int B, int K, int HW,
              float* grad_in, hp_stream_t stream);

/* ---- a6/a7: pseudo labels, materialised.  regda_4.py:76-86 ; regda_7.py:3026-3039, 3188-3201 */
/* y [B,K,H,W] -> gt, gf [B,K,oh,ow] ; centre = (decode(y) >> shift) ; centres (nullable) [B*K,2] int32 */
int hp_pseudo_label(const float* y, int B, int K, int H, int W, int kind, int oh, int ow, int shift,
                    int tmp, const float* tab, float* gt, float* gf, int32_t* centres,
                    hp_stream_t stream);

/* ---- a8-a11: regression disparity, fused (no gt/gf materialised).  regda_4.py:129-143 ;
 *      regda_7.py:3250-3268, 3529-3561, 3609-3632 ; criterion = JointsKLLoss(epsilon) ------ */
/* y [B,K,H,W] (decoded, no gradient) ; y_adv [B,K,oh,ow] ; fused (nullable) [B,K,oh,ow] is the
 * y_adv2 argument of x5/x6 ; outputs as hp_kl_fwd plus centres [B*K,2] int32 and
 * stats [B*K,3] = (logsumexp, sum(target+eps), per-map max M used by the x5/x6 normalisation). */
int hp_regdisp_fwd(const float* y, const float* y_adv, const float* fused, const float* weight,
                   int variant, int mode, float epsilon, int B, int K, int H, int W, int oh, int ow,
                   int shift, int tmp, const float* tab, float* per_map, float* per_sample,
                   float* mean, float* stats, int32_t* centres, void* workspace,
                   hp_stream_t stream);
int hp_regdisp_bwd(const float* y_adv, const float* fused, const float* weight, int variant, int mode,
                   float epsilon, int B, int K, int oh, int ow, int tmp, const float* tab,
                   const int32_t* centres, const float* stats, const float* grad_out, int grad_kind,
                   float* grad_in, hp_stream_t stream);
/* RegressionDisparityx6 mode='max' with its y_adv2 argument given UNFUSED, as the two heads train1.py:410-424 builds it from:
 * fused = a_lo * up64(f_lo) + a_mid * up64(f_mid), f_lo [B,K,16,16], f_mid [B,K,32,32] (nn.Upsample bilinear; train1.py:
 * target5 = 0.5 up(y_adv3) + up(y_adv2) -> a_lo 0.5, a_mid 1).  The fused map is interpolated inside the loss kernel from the
 * staged heads (values bit-identical to hp_fuse_multiscale) and never exists in memory: 37,888 B per map instead of 49,152 +
 * the fusion launch (SURVEY.md 8d, configs[2]).  Only x6 / 'max' / 64x64 label grids; anything else -> HP_ERR_ARG /
 * HP_ERR_SHAPE (materialise with hp_fuse_multiscale and call hp_regdisp_fwd).  Outputs as hp_regdisp_fwd / _bwd. */
int hp_regdisp_fwd_heads(const float* y, const float* y_adv, const float* f_lo, int hl, int wl, float a_lo,
                         const float* f_mid, int hm, int wm, float a_mid, const float* weight, int variant,
                         int mode, float epsilon, int B, int K, int H, int W, int oh, int ow, int tmp,
                         const float* tab, float* per_map, float* per_sample, float* mean, float* stats,
                         int32_t* centres, void* workspace, hp_stream_t stream);
int hp_regdisp_bwd_heads(const float* y_adv, const float* f_lo, int hl, int wl, float a_lo, const float* f_mid,
                         int hm, int wm, float a_mid, const float* weight, int variant, int mode, float epsilon,
                         int B, int K, int oh, int ow, int tmp, const float* tab, const int32_t* centres,
                         const float* stats, const float* grad_out, int grad_kind, float* grad_in,
                         hp_stream_t stream);
/* the .ground_truth / .ground_false attributes the reference classes expose, on demand */
int hp_regdisp_materialize(const float* fused, int variant, int B, int K, int oh, int ow, int tmp,
                           const float* tab, const int32_t* centres, float* gt, float* gf,
                           hp_stream_t stream);

/* ---- a12: multiscale fusion.  train1.py:410-424 (nn.Upsample bilinear, align_corners=False) */
/* out[n,H,W] = a_lo * up(lo[n,hl,wl]) + a_mid * up(mid[n,hm,wm]) + a_hi * hi[n,H,W] ;
 * mid and hi are nullable.  target5 = (0.5, 1, -) ; target0 = (1, -, -) ; 3-scale = (0.5, 1, 1). */
int hp_fuse_multiscale(const float* lo, int hl, int wl, float a_lo,
                       const float* mid, int hm, int wm, float a_mid,
                       const float* hi, float a_hi, int n_maps, int H, int W, float* out,
                       hp_stream_t stream);
/* train1.py:410-424 in one call: out [n_maps, H, W] = a_lo * up(lo) + a_mid * up(mid) (`target5`) and
 * out2 [n_maps, H2, W2] = a_lo2 * up(lo) (`target0 = up32(y_adv3)`), both nn.Upsample(mode='bilinear') (align_corners=False).
 * One launch for the driver's geometry (lo x4 and mid x2 -> H x W, H2 x W2 = H/2 x W/2, e.g. 16 / 32 -> 64 and 32); any other
 * geometry: two hp_fuse_multiscale launches.  Results are bit-identical to the two separate calls. */
int hp_fuse_multiscale_pair(const float* lo, int hl, int wl, float a_lo, const float* mid, int hm, int wm, float a_mid,
                            int n_maps, int H, int W, float* out, float a_lo2, int H2, int W2, float* out2,
                            hp_stream_t stream);
/* fuse (as above, never written to memory) + decode + PCK against target coordinates:
 * BASELINE.json configs[3].  tgt_xy [n_maps,2] ; outputs as hp_accuracy. */
int hp_fuse_decode_pck(const float* lo, int hl, int wl, float a_lo,
                       const float* mid, int hm, int wm, float a_mid,
                       const float* hi, float a_hi, const float* tgt_xy, int B, int K, int H, int W,
                       double thr, float* pred_xy, float* maxvals, int32_t* counts, double* acc_out,
                       void* workspace, hp_stream_t stream);

/* ---- the benchmarked pipeline: gen + loss + decode + PCK in ONE pass over `pred` -------
 * replaces generate_target xB (util.py:9-68) + JointsMSELoss (loss.py:55-65) + JointsKLLoss
 * (loss.py:145-158) + accuracy (keypoint_detection.py:63-92); the target is generated in
 * registers and never written.
 *   pred [B,K,H,W] ; joints float64 [B*K,2] ; vis float32 [B*K]
 *   pred_xy [B*K,2], maxvals [B*K], weight_out [B*K]           (per-map outputs)
 *   partial int64 [4 + 2K + 6] = { mse_fx, kl_fx, n_maps, n_elems, hits[K], valid[K],
 *                                  mse_nan, mse_pinf, mse_ninf, kl_nan, kl_pinf, kl_ninf }
 *       loss sums are fixed point (value * 2^40): integer sums are associative, so any block
 *       schedule, slab split or GPU sharding gives bit-identical results.
 *       (+= semantics when accumulate != 0: batch shards / ranks add up, then finalise)
 *   result float64 [4 + K] = { mse, kl, avg_acc, cnt, acc[K] }  (nullable: skip finalise)
 */
int hp_pipeline_fused(const float* pred, const double* joints, const float* vis,
                      int B, int K, int H, int W, double stride_x, double stride_y, int tmp,
                      const float* tab, float kl_epsilon, double thr, int loss_mask,
                      float* pred_xy, float* maxvals, float* weight_out,
                      int64_t* partial, int accumulate, double* result, void* workspace,
                      hp_stream_t stream);
/* Same, with launch flags.
 *   HP_PIPE_OVERLAP_PREV: this launch belongs to a train of launches over independent, already resident
 *   batches (back-to-back calls of this function on one stream).  It is issued as a programmatic dependent
 *   launch on 1/depth of the kernel's block slots, so that `depth` consecutive launches are resident at
 *   once and one launch's start-up and drain overlap the streaming of its neighbours.  Its blocks read
 *   pred / joints / vis / tab as soon as they get a slot, keep their per-map outputs in shared memory, and
 *   write nothing (outputs, partial, result, workspace) until the previous kernel on the stream has
 *   completed - results are bit-identical to serialised launches.
 *   Contract: pred, joints, vis and tab must NOT be produced by the kernel launched right before this one
 *   on `stream`; consecutive launches of a train use different output buffers.
 *   HP_PIPE_DEPTH(d), d = 1..8: launches resident at once (1 = hand-over only); 0/absent = library default.
 *   Without HP_PIPE_OVERLAP_PREV a launch is fully serialised: it reads nothing and writes nothing before the
 *   previous kernel on the stream has completed (it still carries the programmatic-launch attribute, so that block
 *   scheduling and barrier set-up hide the launch gap behind the previous kernel's tail - nothing else overlaps).
 *   HP_PIPE_DEFER_EXCHANGE (sharded steps only: hp_pipeline_fused_peer / plans with world > 1): the step sends its
 *   partial vector to the other ranks but does not wait for theirs; `partial` / `result` of step s are completed by
 *   the next sharded step on the same workspace (or by hp_pipeline_flush_peer after the last one) - nothing on the
 *   step path then waits for a peer that is less than a whole step late.  Keep step s's output buffers alive and
 *   unread until then. */
#define HP_PIPE_OVERLAP_PREV 1u
#define HP_PIPE_DEFER_EXCHANGE 2u
#define HP_PIPE_DEPTH(d) (((unsigned int)(d) & 15u) << 8)
int hp_pipeline_fused_ex(const float* pred, const double* joints, const float* vis,
                         int B, int K, int H, int W, double stride_x, double stride_y, int tmp,
                         const float* tab, float kl_epsilon, double thr, int loss_mask,
                         float* pred_xy, float* maxvals, float* weight_out,
                         int64_t* partial, int accumulate, double* result, void* workspace,
                         unsigned int flags, hp_stream_t stream);
/* Profiling aid (not part of the reference-facing path): when a buffer is set, the blocks of the TMA-staged
 * pipeline kernel stamp their timeline (entry/exit, and per warp and map: wait begins, data landed, refill
 * issued, map closed) into it; launches alternate between two slots of hp_debug_pipeline_trace_words()
 * uint64 words.  buf = NULL switches it off.  Read by profiles/trace_pipeline.py. */
size_t hp_debug_pipeline_trace_words(void);
int hp_debug_pipeline_trace(void* buf, size_t words);
/* The same aid for the dense 'max' disparity kernel (csrc/hp_regdisp_dense.cuh): per block entry / exit, per consumer warp
 * and map {wait begins, data landed, per-sample label seen, map done}, per sample the builder's {begin, published}
 * (globaltimer ns).  Read by profiles/trace_regdisp.py. */
size_t hp_debug_regdisp_trace_words(void);
int hp_debug_regdisp_trace(void* buf, size_t words);
/* ---- row f3: the loss weightings JointsMSELoss0 / JointsKLLoss5 (uda/model/loss.py:68-112, :160-216) ----
 * hp_mse0_*: both maps shifted by 1e-7 and normalised to sum 1 per map, then 0.5*w*(p-t)^2; per_map [B*K] = mean over HW
 *   ('none'), *mean (nullable) = mean over all elements.  grad_kind HP_GRAD_SCALAR ('mean') / HP_GRAD_PER_MAP ('none').
 * hp_kl5_*: per-map scale s = w3 / max(w3), w3 = sum (out/max(out)) (tgt/max(tgt)) (global maxima, no gradient), then the KL
 *   loss of s*out against s*tgt without target weights; per_sample [B] = mean over K ('none'), *mean = mean over B*K.
 *   scratch: float [4*B*K] (the per-map scale lands at scratch + 3*B*K and is what hp_kl5_bwd takes); stats [B*K,2]. */
int hp_mse0_fwd(const float* out, const float* tgt, const float* weight /*nullable*/, int B, int K, int HW,
                float* per_map, float* mean /*nullable*/, void* workspace, hp_stream_t stream);
int hp_mse0_bwd(const float* out, const float* tgt, const float* weight /*nullable*/, const float* grad_out, int grad_kind,
                int B, int K, int HW, float* grad_in, hp_stream_t stream);
int hp_kl5_fwd(const float* out, const float* tgt, float epsilon, int B, int K, int HW, float* scratch,
               float* per_map, float* per_sample /*nullable*/, float* mean /*nullable*/, float* stats, void* workspace,
               hp_stream_t stream);
int hp_kl5_bwd(const float* out, const float* tgt, float epsilon, const float* scale, const float* stats,
               const float* grad_out, int grad_kind, int B, int K, int HW, float* grad_in, hp_stream_t stream);

/* ---- row f3: the label-fusing disparity variants RegressionDisparity2/3/5/6/7/8 (uda/model/regda_4.py:145-645) ----
 * centres_* int32 [B*K,2]: decoded pseudo-label centres of y, label_1, label_2 (hp_argmax_decode / hp_pseudo_label);
 * writes gt [B,K,H,W] (Gaussian at centres_y), gf = clip(label_p - 10 gt) and/or the per-sample label_p [B,H,W]
 * (any of the three may be NULL).  One block per sample, H*W <= 4096. */
#define HP_LF_RD2 2
#define HP_LF_RD3 3
#define HP_LF_RD5 5
#define HP_LF_RD6 6
#define HP_LF_RD7 7
#define HP_LF_RD8 8
int hp_label_fusion(const int32_t* centres_y, const int32_t* centres_1, const int32_t* centres_2 /*nullable*/,
                    int rule, int B, int K, int H, int W, int tmp, const float* tab,
                    float* gt, float* gf, float* label_p, hp_stream_t stream);

/* ---- small boundary operators around the decode (csrc/hp_extras.cu) ----
 * hp_argmax_decode_f64: get_max_preds for float64 heatmaps (utils/keypoint_detection.py:7-35 takes any ndarray dtype);
 *   preds float32 [n_maps,2], maxvals float64 [n_maps] (the reference returns maxvals in the input dtype).
 * hp_refine_quarter: OPT-IN quarter-pixel refinement of decoded coordinates, in place (not in the reference - SURVEY.md
 *   row a13): preds += 0.25 * sign(hm[y][x+1]-hm[y][x-1], hm[y+1][x]-hm[y-1][x]) for maxima with 1 < x < W-1, 1 < y < H-1.
 * hp_group_accuracy: uda/dataset/keypoint_dataset.py:58-71 on the device: out[g] = mean of acc[index[offsets[g]:offsets[g+1]]],
 *   summed left to right like Python's sum() (bit-equal float64). */
int hp_argmax_decode_f64(const double* heat, int n_maps, int H, int W, float* preds, double* maxvals, hp_stream_t stream);
int hp_refine_quarter(const float* heat, int n_maps, int H, int W, float* preds, hp_stream_t stream);
int hp_group_accuracy(const double* acc, const int32_t* offsets, const int32_t* index, int n_groups, double* out,
                      hp_stream_t stream);
/* partial (e.g. after an NCCL all-reduce over ranks) -> result, on device */
int hp_pipeline_finalize(const int64_t* partial, int K, double* result, hp_stream_t stream);

/* ---- the path's single collective over NVLink peer memory (no NCCL on the per-step path) ----
 * Setup (once; the only entry points that own memory): every rank allocates a mailbox, exports its
 * 64-byte CUDA IPC handle, exchanges handles with the other ranks of the node (any transport) and maps
 * theirs.  Per step: hp_pipeline_finalize_peer replaces [all-reduce(partial) ; hp_pipeline_finalize]:
 * it writes this rank's partial into every mailbox, waits (bounded) for all ranks' step `seq`, sums in
 * rank order and finalises.  mailboxes[r] = rank r's mailbox as mapped here (own pointer for r == rank).
 * `seq` = 1, 2, 3, ... identical on all ranks, or 0: the step number is counted on the device (in the rank's own
 * mailbox), which makes a captured CUDA graph of steps replayable.  Only 0 is accepted by this version.
 * On a timeout (a rank did not deliver within ~2 s) every entry of `result` is NaN and every entry of the reduced
 * partial vector is -1; the step is not counted.
 * hp_pipeline_flush_peer completes the outstanding step of a train of HP_PIPE_DEFER_EXCHANGE steps that used
 * `workspace`; a no-op when nothing is outstanding.
 * hp_pck_finalize_peer is the same exchange for configs[3]: counts int32 [2K] (hits, valid) of this rank -> totals in
 * counts_out and acc_out[K+2] = acc[K], avg_acc, cnt (replaces [all-reduce(counts) ; hp_pck_finalize]). */
size_t hp_peer_mailbox_bytes(int world);
int hp_peer_alloc(int world, void** mailbox);
int hp_peer_free(void* mailbox);
int hp_peer_export(void* mailbox, void* handle64);
int hp_peer_import(const void* handle64, void** mapped);
int hp_peer_close(void* mapped);
int hp_pipeline_finalize_peer(const int64_t* partial, void* const* mailboxes, int rank, int world, int K,
                              int64_t seq, int64_t* partial_out, double* result, hp_stream_t stream);
int hp_pipeline_flush_peer(void* workspace, void* const* mailboxes, int rank, int world, hp_stream_t stream);
int hp_pck_finalize_peer(const int32_t* counts, void* const* mailboxes, int rank, int world, int K,
                         int32_t* counts_out, double* acc_out, hp_stream_t stream);
/* hp_fuse_decode_pck on this rank's samples with the exchange folded in: `counts` / `acc_out` hold the totals over ALL ranks
 * (keypoint_detection.py:63-92 on the concatenated batch).  For the shapes of the staged kernel (32 / 64 / 128) the kernel's
 * last block sums the 2K integer counts over the mailboxes itself - one launch per step; for any other geometry a one-warp
 * hp_pck_finalize_peer launch follows.  world == 1: plain hp_fuse_decode_pck.
 * flags & HP_PIPE_DEFER_EXCHANGE (a step of a train; staged kernel only): the step only SENDS its counts; its totals land in
 * partial_out (int64 [4+2K+6], the pipeline's vector: hits / valid at [4, 4+2K)) and result_out (double [4+K]:
 * -, -, avg_acc, cnt, acc[K]) when the next such step on the same workspace - or hp_pipeline_flush_peer - collects them;
 * `counts` / `acc_out` are not written then. */
int hp_fuse_decode_pck_peer(const float* lo, int hl, int wl, float a_lo, const float* mid, int hm, int wm, float a_mid,
                            const float* hi, float a_hi, const float* tgt_xy, int B, int K, int H, int W, double thr,
                            float* pred_xy, float* maxvals, int32_t* counts, double* acc_out, void* workspace,
                            void* const* mailboxes, int rank, int world, unsigned int flags, int64_t* partial_out,
                            double* result_out, hp_stream_t stream);

/* One batch-sharded step in a single call and - for the shapes served by the TMA-staged kernel (H*W = 256,
 * 1024 or a multiple of 4096 floats, aligned) - in a single KERNEL: the last block of the fused kernel stores
 * this rank's partial vector into every rank's mailbox over NVLink, waits (bounded) for the others, sums in rank
 * order and finalises; `partial` and `result` hold the totals over all ranks on return (stream order).  Other
 * shapes run hp_pipeline_fused_ex + hp_pipeline_finalize_peer.  `seq` must be 0 (step counted on the device).
 * With HP_PIPE_OVERLAP_PREV a train of sharded steps stays overlapped. */
int hp_pipeline_fused_peer(const float* pred, const double* joints, const float* vis,
                           int B, int K, int H, int W, double stride_x, double stride_y, int tmp,
                           const float* tab, float kl_epsilon, double thr, int loss_mask,
                           float* pred_xy, float* maxvals, float* weight_out,
                           int64_t* partial, double* result, void* workspace,
                           void* const* mailboxes, int rank, int world, int64_t seq,
                           unsigned int flags, hp_stream_t stream);

/* Pre-bound steps.  A step is a ~12 us kernel; marshalling two dozen arguments through an FFI costs about as much
 * on the host, so a plan validates and stores the arguments of hp_pipeline_fused_ex (world <= 1, mailboxes NULL) or
 * hp_pipeline_fused_peer (world > 1) once and is launched with a two-argument call.  The plan keeps the caller's
 * pointers (it owns none of the memory) until hp_pipeline_plan_destroy. */
typedef struct hp_plan hp_plan_t;
int hp_pipeline_plan_create(const float* pred, const double* joints, const float* vis,
                            int B, int K, int H, int W, double stride_x, double stride_y, int tmp,
                            const float* tab, float kl_epsilon, double thr, int loss_mask,
                            float* pred_xy, float* maxvals, float* weight_out,
                            int64_t* partial, int accumulate, double* result, void* workspace,
                            void* const* mailboxes, int rank, int world, unsigned int flags,
                            hp_plan_t** plan);
int hp_pipeline_plan_launch(const hp_plan_t* plan, hp_stream_t stream);
int hp_pipeline_plan_destroy(hp_plan_t* plan);

/* Host-buffer form (end-to-end path): h_* are pinned host arrays; the batch is cut into slabs
 * of slab_B samples whose H2D copies (copy_stream) overlap the kernels (stream); device
 * staging d_pred holds 2 slabs [2*slab_B,K,H,W]; returns after h_result is valid.
 * Sharded (world > 1, mailboxes as for hp_pipeline_fused_peer; else NULL, 0, 1): h_pred .. are THIS rank's slice; after
 * the last slab the ranks' partial vectors are exchanged over the peer mailboxes, so d_partial / h_result hold the
 * totals over all ranks (h_pred_xy stays this rank's slice). */
int hp_pipeline_fused_host(const float* h_pred, const double* h_joints, const float* h_vis,
                           int B, int K, int H, int W, double stride_x, double stride_y, int tmp,
                           const float* tab, float kl_epsilon, double thr, int loss_mask,
                           int slab_B, float* d_pred, double* d_joints, float* d_vis,
                           float* d_pred_xy, float* d_maxvals, float* d_weight,
                           int64_t* d_partial, double* d_result, void* workspace,
                           float* h_pred_xy /*nullable*/, double* h_result,
                           void* const* mailboxes /*nullable*/, int rank, int world,
                           hp_stream_t stream, hp_stream_t copy_stream);

#ifdef __cplusplus
}
#endif
#endif /* HP_B200_H */
