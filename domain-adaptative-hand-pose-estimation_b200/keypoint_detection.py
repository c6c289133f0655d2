"""Decode + PCK on the GPU behind the reference's signatures.

Mirror of ``utils/keypoint_detection.py:7-92`` (``get_max_preds``, ``accuracy``).  The reference
functions take **numpy** arrays (``train1.py:464-475`` passes ``y.detach().cpu().numpy()``); these
take numpy arrays *or* CUDA tensors.  numpy in -> numpy out (same dtypes/shapes as the reference);
CUDA tensor in -> tensors out with no host hop for the heatmaps.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError("no CUDA device: the B200 heatmap path has no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _as_cuda_heatmaps(x, what):
    """-> (cuda contiguous tensor, was_numpy, numpy dtype or None).  float32 / float16 / bfloat16 become float32 (exact
    widening); float64 stays float64 and is decoded by the float64 kernel - a cast to float32 could merge two distinct
    maxima into a tie and move the argmax (utils/keypoint_detection.py:12-21 accepts any ndarray dtype)."""
    if isinstance(x, np.ndarray):
        if x.dtype not in (np.float32, np.float16, np.float64):
            raise TypeError(f"{what}: unsupported dtype {x.dtype}")
        t = torch.from_numpy(np.ascontiguousarray(x)).to(_device(), non_blocking=False)
        if x.dtype == np.float64:
            return t.contiguous(), True, x.dtype
        return _lib.require_cuda(t, what), True, x.dtype
    if isinstance(x, torch.Tensor):
        if x.dtype == torch.float64:
            if not x.is_cuda:
                raise RuntimeError(f"{what}: tensor is on {x.device}; the B200 heatmap path runs on CUDA only "
                                   "(there is no CPU fallback)")
            return x.detach().contiguous(), False, None
        return _lib.require_cuda(x.detach(), what), False, None
    raise AssertionError("batch_heatmaps should be numpy.ndarray")  # keypoint_detection.py:12-13


def decode(heat: torch.Tensor, refine=None):
    """CUDA tensor [B,K,H,W] (float32, or float64) -> (preds float32 [B,K,2] as (x,y), maxvals [B,K,1] in the input's
    precision) on device.  utils/keypoint_detection.py:16-35 semantics (first index on ties, NaN wins, masked by max>0).

    ``refine="quarter"`` (opt-in; NOT in the reference, whose ``get_max_preds`` has no refinement - SURVEY.md row a13):
    shifts each interior maximum by a quarter pixel towards its higher neighbour,
    ``coord += 0.25 * sign(hm[y][x+1] - hm[y][x-1])`` (resp. y) for ``1 < x < W-1, 1 < y < H-1``."""
    if refine not in (None, "quarter"):
        raise ValueError("refine must be None or 'quarter'")
    f64 = isinstance(heat, torch.Tensor) and heat.dtype == torch.float64
    if f64:
        if not heat.is_cuda:
            raise RuntimeError("decode: tensor is on cpu; the B200 heatmap path runs on CUDA only (there is no CPU fallback)")
        heat = heat.contiguous()
    else:
        heat = _lib.require_cuda(heat, "decode")
    B, K, H, W = heat.shape
    if H * W == 0:
        raise ValueError("attempt to get argmax of an empty sequence")
    preds = torch.empty((B, K, 2), dtype=torch.float32, device=heat.device)
    maxvals = torch.empty((B, K, 1), dtype=heat.dtype if f64 else torch.float32, device=heat.device)
    with _lib.on_device(heat.device):
        st = _lib.stream_ptr(heat.device)
        if f64:
            _lib.call("hp_argmax_decode_f64", _lib.ptr(heat), B * K, H, W, _lib.ptr(preds), _lib.ptr(maxvals), st)
            if refine == "quarter":
                _lib.call("hp_refine_quarter", _lib.ptr(heat.float()), B * K, H, W, _lib.ptr(preds), st)
        else:
            _lib.call("hp_argmax_decode", _lib.ptr(heat), B * K, H, W, _lib.ptr(preds), _lib.ptr(maxvals), None, st)
            if refine == "quarter":
                _lib.call("hp_refine_quarter", _lib.ptr(heat), B * K, H, W, _lib.ptr(preds), st)
    return preds, maxvals


def get_max_preds(batch_heatmaps, refine=None):
    """utils/keypoint_detection.py:7-35.  numpy [B,K,H,W] -> (preds float32 [B,K,2], maxvals [B,K,1] in the input dtype).
    ``refine``: see :func:`decode` (opt-in extension, default off = the reference)."""
    if not isinstance(batch_heatmaps, (np.ndarray, torch.Tensor)):
        raise AssertionError("batch_heatmaps should be numpy.ndarray")
    assert batch_heatmaps.ndim == 4, "batch_images should be 4-ndim"
    heat, was_numpy, np_dtype = _as_cuda_heatmaps(batch_heatmaps, "get_max_preds")
    preds, maxvals = decode(heat, refine)
    if was_numpy:
        return preds.cpu().numpy(), maxvals.cpu().numpy().astype(np_dtype, copy=False)
    return preds, maxvals


def group_accuracy(accuracies, keypoints_group):
    """uda/dataset/keypoint_dataset.py:58-71 on the device: per-group means of the per-joint accuracies.
    ``accuracies``: float64 [K] CUDA tensor (e.g. ``pck(...)[0][:K]``) or anything ``torch.as_tensor`` takes;
    ``keypoints_group``: ``{name: indices}``.  -> ``{name: float}`` (one small device->host read), summed left to right
    like the reference's ``sum(...) / len(...)`` so the float64 values are bit-equal."""
    acc = torch.as_tensor(accuracies, dtype=torch.float64)
    if not acc.is_cuda:
        acc = acc.to(_device())
    acc = acc.contiguous()
    names = list(keypoints_group)
    offsets, index = [0], []
    for n in names:
        idx = [int(i) for i in keypoints_group[n]]
        if not idx:
            raise ZeroDivisionError("division by zero")          # what the reference's sum(...)/len(...) raises
        if max(idx) >= acc.numel() or min(idx) < -acc.numel():
            raise IndexError("index out of range")
        index += [i % acc.numel() for i in idx]
        offsets.append(len(index))
    dev = acc.device
    off_t = torch.tensor(offsets, dtype=torch.int32, device=dev)
    idx_t = torch.tensor(index, dtype=torch.int32, device=dev)
    out = torch.empty((len(names),), dtype=torch.float64, device=dev)
    with _lib.on_device(dev):
        _lib.call("hp_group_accuracy", _lib.ptr(acc), _lib.ptr(off_t), _lib.ptr(idx_t), len(names), _lib.ptr(out),
                  _lib.stream_ptr(dev))
    host = out.cpu().numpy()
    return {n: float(host[i]) for i, n in enumerate(names)}


def pck(output: torch.Tensor, target: torch.Tensor, thr: float = 0.5):
    """Device-side accuracy: -> (acc_vec float64 [K+2] = acc[K], avg_acc, cnt ; pred_xy float32 [B,K,2] ;
    counts int32 [2K] = hits, valid), all CUDA tensors, one launch, no synchronisation."""
    output = _lib.require_cuda(output, "accuracy(output)")
    target = _lib.require_cuda(target, "accuracy(target)")
    if output.shape != target.shape or output.ndim != 4:
        raise ValueError(f"accuracy: output {tuple(output.shape)} vs target {tuple(target.shape)}")
    B, K, H, W = output.shape
    if K > _lib.MAX_K:
        raise ValueError(f"accuracy: K={K} exceeds HP_MAX_K={_lib.MAX_K}")
    dev = output.device
    pred_xy = torch.empty((B, K, 2), dtype=torch.float32, device=dev)
    counts = torch.empty((2 * K,), dtype=torch.int32, device=dev)
    acc = torch.empty((K + 2,), dtype=torch.float64, device=dev)
    with _lib.on_device(dev):
        ws = _lib.workspace(dev, B * K, K)
        _lib.call("hp_accuracy", _lib.ptr(output), _lib.ptr(target), B, K, H, W, C.c_double(thr), _lib.ptr(pred_xy),
                  _lib.ptr(counts), _lib.ptr(acc), _lib.ptr(ws), _lib.stream_ptr(dev))
    return acc, pred_xy, counts


def _pck_from_decodes(output, target, thr):
    """accuracy() for inputs the fused kernel does not take (float64 heatmaps): decode both, then the PCK count and
    finalise kernels on the coordinates.  Same integer counts / float64 ratios."""
    if output.shape != target.shape or output.ndim != 4:
        raise ValueError(f"accuracy: output {tuple(output.shape)} vs target {tuple(target.shape)}")
    B, K, H, W = output.shape
    if K > _lib.MAX_K:
        raise ValueError(f"accuracy: K={K} exceeds HP_MAX_K={_lib.MAX_K}")
    pred_xy, _ = decode(output)
    tgt_xy, _ = decode(target)
    dev = output.device
    counts = torch.zeros((2 * K,), dtype=torch.int32, device=dev)
    acc = torch.empty((K + 2,), dtype=torch.float64, device=dev)
    with _lib.on_device(dev):
        _lib.call("hp_pck_accumulate", _lib.ptr(pred_xy), _lib.ptr(tgt_xy), B, K, H, W, C.c_double(thr), _lib.ptr(counts),
                  _lib.stream_ptr(dev))
        _lib.call("hp_pck_finalize", _lib.ptr(counts), K, _lib.ptr(acc), _lib.stream_ptr(dev))
    return acc, pred_xy


def accuracy(output, target, hm_type="gaussian", thr=0.5):
    """utils/keypoint_detection.py:63-92.  -> (acc float64[K], avg_acc, cnt, pred [B,K,2]).

    ``pred`` is a numpy array when ``output`` was one, else a CUDA tensor; ``acc``/``avg_acc``/``cnt``
    are host values like the reference's (one 8*(K+2)-byte device->host read)."""
    if hm_type != "gaussian":
        # the reference leaves `pred` undefined in this case (keypoint_detection.py:72-78)
        raise NameError("accuracy: only hm_type='gaussian' is defined by the reference")
    out_t, was_numpy, _ = _as_cuda_heatmaps(output, "accuracy(output)")
    tgt_t, _, _ = _as_cuda_heatmaps(target, "accuracy(target)")
    if out_t.dtype == torch.float64 or tgt_t.dtype == torch.float64:
        acc_vec, pred_xy = _pck_from_decodes(out_t, tgt_t, thr)     # float64 heatmaps: two decodes + the count kernels
    else:
        acc_vec, pred_xy, _ = pck(out_t, tgt_t, thr)
    host = acc_vec.cpu().numpy()
    K = out_t.shape[1]
    acc = host[:K].copy()
    cnt = int(host[K + 1])
    avg_acc = float(host[K]) if cnt != 0 else 0
    pred = pred_xy.cpu().numpy() if was_numpy else pred_xy
    return acc, avg_acc, cnt, pred


# ------------------------------------------------------------------------------------------
# the decode helpers next to get_max_preds (utils/keypoint_detection.py:139-239)
# ------------------------------------------------------------------------------------------

def _argmax_idx(heat):
    """CUDA [N,K,H,W] -> (flat argmax int32 [N*K], maxvals float32 [N*K]) with the decode kernel's rules
    (first index on ties, NaN wins - what torch.max / torch.argmax do on the reference's side)."""
    heat = _lib.require_cuda(heat, "argmax")
    B, K, H, W = heat.shape
    dev = heat.device
    preds = torch.empty((B * K, 2), dtype=torch.float32, device=dev)
    maxvals = torch.empty((B * K,), dtype=torch.float32, device=dev)
    idx = torch.empty((B * K,), dtype=torch.int32, device=dev)
    with _lib.on_device(dev):
        _lib.call("hp_argmax_decode", _lib.ptr(heat), B * K, H, W, _lib.ptr(preds), _lib.ptr(maxvals), _lib.ptr(idx),
                  _lib.stream_ptr(dev))
    return idx, maxvals


def find_keypoints_max(heatmaps):
    """utils/keypoint_detection.py:139-154.  [C,H,W] -> [C,3] = (u, v, max) with the reference's index arithmetic
    (``v = floor(ind / size(1))``, ``u = fmod(ind, size(2))``)."""
    hm = _lib.require_cuda(heatmaps.detach(), "find_keypoints_max")
    if hm.ndim != 3:
        raise ValueError("find_keypoints_max: heatmaps must be [C,H,W]")
    C_, H, W = hm.shape
    idx, maxvals = _argmax_idx(hm.reshape(1, C_, H, W))
    ind = idx.to(torch.float32)
    v = torch.floor(torch.div(ind, H))
    u = torch.fmod(ind, W)
    return torch.cat((u.view(-1, 1), v.view(-1, 1), maxvals.view(-1, 1)), 1)


def _resized(hm, resize_dim):
    from .fusion import upsample_bilinear
    size = (resize_dim, resize_dim) if isinstance(resize_dim, int) else (int(resize_dim[0]), int(resize_dim[1]))
    return upsample_bilinear(hm, size), size


def compute_uv_from_heatmaps(hm, resize_dim):
    """utils/keypoint_detection.py:156-171: bilinear resize (CUDA fusion kernel), argmax -> float [B,K,2] (u, v),
    NOT masked by the maximum."""
    resized, (H, W) = _resized(hm, resize_dim)
    B, K = resized.shape[:2]
    uvc = find_keypoints_max(resized.view(-1, H, W))
    return uvc.view(-1, K, 3)[:, :, :2]


_MAX_K = 64   # HP_MAX_K of include/hp_b200.h (per-joint counters of the fused decode kernel)


def compute_uv_from_heatmaps2(hm, resize_dim):
    """utils/keypoint_detection.py:174-205: bilinear resize, argmax -> int64 [B,K,2] (x, y) zeroed where max <= 0.

    One launch: the resized map is decoded in registers and never written (``hp_fuse_decode_pck`` with a single
    source and weight 1 - the arithmetic of :func:`fusion.upsample_bilinear`, so the same map bit for bit); the
    kernel's coordinates are already ``idx % W``, ``idx // W`` masked by ``max > 0``."""
    hm = _lib.require_cuda(hm.detach(), "compute_uv_from_heatmaps2")
    size = (resize_dim, resize_dim) if isinstance(resize_dim, int) else (int(resize_dim[0]), int(resize_dim[1]))
    B, K = hm.shape[:2]
    H, W = size
    if K > _MAX_K:   # the two-launch composition (materialised resize, then the decode kernel)
        resized, _ = _resized(hm, resize_dim)
        idx, maxvals = _argmax_idx(resized)
        idx = idx.to(torch.int64).view(B, K, 1)
        preds = idx.repeat(1, 1, 2)
        preds[:, :, 0] = preds[:, :, 0] % W
        preds[:, :, 1] = torch.div(preds[:, :, 1], W, rounding_mode="floor")
        preds *= torch.greater(maxvals.view(B, K, 1), 0.0).repeat(1, 1, 2)
        return preds
    dev = hm.device
    pred_xy = torch.empty((B, K, 2), dtype=torch.float32, device=dev)
    maxvals = torch.empty((B, K, 1), dtype=torch.float32, device=dev)
    tgt = torch.zeros((B * K, 2), dtype=torch.float32, device=dev)      # the kernel also scores PCK; unused here
    acc = torch.empty((K + 2,), dtype=torch.float64, device=dev)
    with _lib.on_device(dev):
        ws = _lib.workspace(dev, B * K, K)
        _lib.call("hp_fuse_decode_pck", _lib.ptr(hm), hm.shape[2], hm.shape[3], C.c_float(1.0), None, 0, 0, C.c_float(0.0),
                  None, C.c_float(0.0), _lib.ptr(tgt), B, K, H, W, C.c_double(0.5), _lib.ptr(pred_xy), _lib.ptr(maxvals),
                  None, _lib.ptr(acc), _lib.ptr(ws), _lib.stream_ptr(dev))
    return pred_xy.to(torch.int64)


def compute_uv_from_heatmaps3(heatmap, beta=100.0, scale=4.0):
    """utils/keypoint_detection.py:209-239: soft-argmax, ``softmax(100*h)``-weighted mean of the pixel coordinates,
    -> float32 [B,K,2] = 4 * (E[col], E[row]).  One CUDA pass, nothing materialised."""
    hm = _lib.require_cuda(heatmap.detach(), "compute_uv_from_heatmaps3")
    B, K, H, W = hm.shape
    out = torch.empty((B, K, 2), dtype=torch.float32, device=hm.device)
    with _lib.on_device(hm.device):
        _lib.call("hp_soft_argmax", _lib.ptr(hm), B * K, H, W, float(beta), float(scale), _lib.ptr(out),
                  _lib.stream_ptr(hm.device))
    return out
