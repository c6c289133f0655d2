"""CPU ORACLE for the heatmap keypoint hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT.

A CPU restatement (numpy + torch-CPU) of the reference algorithm for every row of
SURVEY.md §8(a).  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this module, and only as the
checker / the timed CPU baseline.  The product path (the package next to this directory)
never imports it and fails loudly when the CUDA library is missing.

Parity status: PINNED.  The reference ships no tests or golden vectors, so
the pin is (i) ``tests/test_oracle_vs_reference.py``,
which executes the real reference from ``/root/reference`` (via ``oracle/ref_loader.py``) next
to this restatement on seeded inputs and demands bit-equality, and (ii) the committed
fixtures in ``tests/golden/`` that ``oracle/gen_golden.py`` produced FROM THE REAL REFERENCE.

The restatement keeps the reference's *algorithmic structure* (host numpy argmax, the
W*H*H*W look-up table of Gaussians, per-(sample, joint) Python loops, torch's stock
log_softmax / KLDivLoss / MSELoss / bilinear Upsample), because it doubles as the timed CPU
baseline: its cost profile must be the reference's.  Third-party arithmetic the reference
leans on: torch (``nn.MSELoss``, ``F.log_softmax``, ``nn.KLDivLoss``,
``nn.Upsample``) and numpy (``argmax``, ``amax``, ``exp``, ``dot``, ``linalg.norm``) -
unpinned by the reference (``README.md:11-18``: ``torch>=1.7.0``, numpy unversioned); this
container has torch 2.11.0 / numpy 2.3.5.

All ``file:line`` citations are relative to the reference tree.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

# ----------------------------------------------------------------------------------------
# a1  decode                                            utils/keypoint_detection.py:7-35
# ----------------------------------------------------------------------------------------

def get_max_preds(batch_heatmaps: np.ndarray):
    """Per-map argmax decode.  utils/keypoint_detection.py:7-35.

    Returns ``preds float32[B,K,2]`` as (x, y) and ``maxvals[B,K,1]`` in the input dtype.
    First (lowest) flat index wins ties, NaN counts as the maximum (numpy argmax), and the
    coordinates are zeroed wherever ``maxval > 0`` is false (:30-34).
    """
    if not isinstance(batch_heatmaps, np.ndarray):
        raise AssertionError("batch_heatmaps should be numpy.ndarray")       # :12-13
    if batch_heatmaps.ndim != 4:
        raise AssertionError("batch_images should be 4-ndim")                # :14
    n, k, _, w = batch_heatmaps.shape
    flat = batch_heatmaps.reshape(n, k, -1)
    where = flat.argmax(axis=2)                                              # :20
    peak = flat.max(axis=2).reshape(n, k, 1)                                 # :21-23
    xy = np.empty((n, k, 2), dtype=np.float32)
    as_f32 = where.astype(np.float32)                                        # :26 (index -> fp32)
    xy[:, :, 0] = as_f32 % w                                                 # :28
    xy[:, :, 1] = np.floor(as_f32 / w)                                       # :29
    keep = (peak > 0.0).astype(np.float32)                                   # :31-32
    xy *= keep                                                               # :34 (broadcast == tile)
    return xy, peak


# ----------------------------------------------------------------------------------------
# a2  PCK                                               utils/keypoint_detection.py:38-92
# ----------------------------------------------------------------------------------------

def calc_dists(preds, target, normalize):
    """Normalised float64 distance per (joint, sample); -1 where the target is not > 1 in both
    coordinates.  utils/keypoint_detection.py:38-50 (result layout is [K, B])."""
    preds = preds.astype(np.float32)
    target = target.astype(np.float32)
    n, k = preds.shape[0], preds.shape[1]
    out = np.zeros((k, n))
    for s in range(n):
        for j in range(k):
            if target[s, j, 0] > 1 and target[s, j, 1] > 1:                   # :44
                a = preds[s, j, :] / normalize[s]                            # :45
                b = target[s, j, :] / normalize[s]                           # :46
                out[j, s] = np.linalg.norm(a - b)                            # :47
            else:
                out[j, s] = -1                                               # :49
    return out


def dist_acc(dists, thr=0.5):
    """Fraction of valid (!= -1) distances below ``thr``; -1 when none is valid.
    utils/keypoint_detection.py:53-60."""
    valid = dists != -1
    n_valid = valid.sum()
    if n_valid > 0:
        return (dists[valid] < thr).sum() * 1.0 / n_valid
    return -1


def pck_counts(preds, target, h, w, thr=0.5):
    """Integer (hits[K], valid[K]) behind ``accuracy`` - the quantities that must be bit-exact."""
    norm = np.ones((preds.shape[0], 2)) * np.array([h, w]) / 10
    d = calc_dists(preds, target, norm)
    valid = d != -1
    hits = (d < thr) & valid
    return hits.sum(axis=1).astype(np.int64), valid.sum(axis=1).astype(np.int64)


def accuracy(output, target, hm_type="gaussian", thr=0.5):
    """PCK from two heatmap tensors.  utils/keypoint_detection.py:63-92.

    Returns ``(acc float64[K], avg_acc, cnt, pred float32[B,K,2])``."""
    k = output.shape[1]
    norm = 1.0
    if hm_type == "gaussian":
        pred, _ = get_max_preds(output)                                      # :73
        target, _ = get_max_preds(target)                                    # :74
        h, w = output.shape[2], output.shape[3]
        norm = np.ones((pred.shape[0], 2)) * np.array([h, w]) / 10           # :77
    d = calc_dists(pred, target, norm)                                       # :78
    acc = np.zeros(k)
    total, cnt = 0, 0
    for j in range(k):                                                       # :84-88
        acc[j] = dist_acc(d[j], thr)
        if acc[j] >= 0:
            total = total + acc[j]
            cnt += 1
    avg = total / cnt if cnt != 0 else 0                                     # :90
    return acc, avg, cnt, pred


# ----------------------------------------------------------------------------------------
# Gaussian patch shared by a5/a6/a7                     uda/dataset/util.py:49-54 ; regda_4.py:56-61
# ----------------------------------------------------------------------------------------

def gaussian_patch(sigma, tmp_size):
    """The un-normalised (2*tmp+1)^2 float32 Gaussian, evaluated exactly like the reference
    does (float32 ``arange``, float32 ``np.exp``).  uda/dataset/util.py:49-54."""
    size = 2 * tmp_size + 1
    ax = np.arange(0, size, 1, np.float32)
    ay = ax[:, np.newaxis]
    c = size // 2
    return np.exp(-((ax - c) ** 2 + (ay - c) ** 2) / (2 * sigma ** 2))


def _paste_ranges(mu, tmp_size, extent):
    """(patch_lo, patch_hi, img_lo, img_hi) along one axis.  uda/dataset/util.py:39-40,56-61."""
    ul = int(mu - tmp_size)
    br = int(mu + tmp_size + 1)
    return max(0, -ul), min(br, extent) - ul, max(0, ul), min(br, extent)


# ----------------------------------------------------------------------------------------
# a5  target generation                                 uda/dataset/util.py:9-68
# ----------------------------------------------------------------------------------------

def generate_target(joints, joints_vis, heatmap_size, sigma, image_size):
    """Gaussian target + visibility weight for ONE sample.  uda/dataset/util.py:9-68.

    ``heatmap_size`` is (W, H).  Centre = ``int(joint / stride + 0.5)`` in float64 with
    truncation toward zero (:37-38); an out-of-range centre zeroes the weight (:42-46); the
    patch is pasted only when the (possibly fractional) visibility is > 0.5 (:63-66)."""
    k = joints.shape[0]
    weight = np.ones((k, 1), dtype=np.float32)
    weight[:, 0] = joints_vis[:, 0]                                          # :22-23
    target = np.zeros((k, heatmap_size[1], heatmap_size[0]), dtype=np.float32)
    tmp_size = sigma * 3                                                     # :30
    image_size = np.array(image_size)
    heatmap_size = np.array(heatmap_size)
    for j in range(k):
        stride = image_size / heatmap_size                                   # :36
        mu_x = int(joints[j][0] / stride[0] + 0.5)                           # :37
        mu_y = int(joints[j][1] / stride[1] + 0.5)                           # :38
        if mu_x >= heatmap_size[0] or mu_y >= heatmap_size[1] or mu_x < 0 or mu_y < 0:
            weight[j] = 0                                                    # :42-46
            continue
        g = gaussian_patch(sigma, tmp_size)                                  # :49-54
        gx0, gx1, ix0, ix1 = _paste_ranges(mu_x, tmp_size, heatmap_size[0])
        gy0, gy1, iy0, iy1 = _paste_ranges(mu_y, tmp_size, heatmap_size[1])
        if weight[j] > 0.5:                                                  # :63-64
            target[j][iy0:iy1, ix0:ix1] = g[gy0:gy1, gx0:gx1]                # :65-66
    return target, weight


def generate_target_batch(joints, joints_vis, heatmap_size, sigma, image_size):
    """B samples through :func:`generate_target`, the way the DataLoader collates them
    (uda/dataset/hand_3d_studio.py:98-104).  joints [B,K,2], joints_vis [B,K,1]."""
    ts, ws = [], []
    for b in range(joints.shape[0]):
        t, w = generate_target(joints[b], joints_vis[b], heatmap_size, sigma, image_size)
        ts.append(t)
        ws.append(w)
    return np.stack(ts), np.stack(ws)


# ----------------------------------------------------------------------------------------
# a3  MSE                                               uda/model/loss.py:27-65
# ----------------------------------------------------------------------------------------

def joints_mse_loss(output, target, target_weight=None, reduction="mean"):
    """0.5*(pred-gt)^2*w; 'mean' over ALL elements, 'none' -> mean over HW -> [B,K].
    uda/model/loss.py:55-65."""
    b, k = output.shape[0], output.shape[1]
    p = output.reshape(b, k, -1)
    g = target.reshape(b, k, -1)
    l = F.mse_loss(p, g, reduction="none") * 0.5                             # :59
    if target_weight is not None:
        l = l * target_weight.view(b, k, 1)                                  # :61
    if reduction == "mean":
        return l.mean()                                                      # :63
    if reduction == "none":
        return l.mean(dim=-1)                                                # :65
    return None


# ----------------------------------------------------------------------------------------
# a4  KL                                                uda/model/loss.py:115-158
# ----------------------------------------------------------------------------------------

def joints_kl_loss(output, target, target_weight=None, reduction="mean", epsilon=0.0):
    """KL(q || softmax(pred)) per map with q = (gt+eps)/sum(gt+eps); 'mean' over B*K,
    'none' -> mean over K -> [B].  uda/model/loss.py:145-158."""
    b, k = output.shape[0], output.shape[1]
    logp = F.log_softmax(output.reshape(b, k, -1), dim=-1)                   # :147-148
    q = target.reshape(b, k, -1) + epsilon                                   # :149-150
    q = q / q.sum(dim=-1, keepdims=True)                                     # :151
    l = F.kl_div(logp, q, reduction="none").sum(dim=-1)                      # :152
    if target_weight is not None:
        l = l * target_weight.view(b, k)                                     # :154
    if reduction == "mean":
        return l.mean()                                                      # :156
    if reduction == "none":
        return l.mean(dim=-1)                                                # :158
    return None


# ----------------------------------------------------------------------------------------
# a6/a7  pseudo labels                                  regda_4.py:17-86 ; regda_7.py:2956-3039, 3118-3201
# ----------------------------------------------------------------------------------------

#: variant -> (output side or None = same as input, divisor applied to the decoded coordinate,
#:             tmp_size as a multiple of sigma)
PLG_VARIANTS = {
    "base": (None, 1, 3),    # PseudoLabelGenerator   regda_4.py:40-48    tmp = 3*sigma   (13x13 @ sigma 2)
    "03":   (32, 2, 2),      # PseudoLabelGenerator03 regda_7.py:3141-3150, 3195  tmp = 2*sigma  (9x9)
    "01":   (16, 4, 1.5),    # PseudoLabelGenerator01 regda_7.py:2979-2988, 3033  tmp = 1.5*sigma (7x7)
}

_LUT_CACHE = {}


def gaussian_lut(width, height, sigma, tmp_size):
    """``lut[mu_x][mu_y]`` = the H x W map with the (truncated) Gaussian pasted at (mu_x, mu_y).
    regda_4.py:46-73 (64 MiB at 64x64)."""
    key = (width, height, sigma, tmp_size)
    if key not in _LUT_CACHE:
        lut = np.zeros((width, height, height, width), dtype=np.float32)
        for mu_x in range(width):
            for mu_y in range(height):
                g = gaussian_patch(sigma, tmp_size)
                gx0, gx1, ix0, ix1 = _paste_ranges(mu_x, tmp_size, width)
                gy0, gy1, iy0, iy1 = _paste_ranges(mu_y, tmp_size, height)
                lut[mu_x][mu_y][iy0:iy1, ix0:ix1] = g[gy0:gy1, gx0:gx1]
        _LUT_CACHE[key] = lut
    return _LUT_CACHE[key]


def pseudo_label(y: torch.Tensor, variant="base", sigma=2):
    """(ground_truth, ground_false) from a prediction ``y [B,K,H,W]``.

    base: regda_4.py:76-86  - centres = int(decode(y)); gf_k = clip(sum_{j!=k} gt_j, 0, 1)
          through the (B,HW,K)x(K,K) float32 dot with 1-I (:83-84).
    03:   regda_7.py:3188-3201 - centres = int(decode/2) on 32x32; gf = clip(1-10*gt).
    01:   regda_7.py:3026-3039 - centres = int(decode/4) on 16x16; gf = clip(1-10*gt)."""
    side, div, tmp_factor = PLG_VARIANTS[variant]
    b, k, h, w = y.shape
    oh, ow = (h, w) if side is None else (side, side)
    lut = gaussian_lut(ow, oh, sigma, sigma * tmp_factor)
    xy, _ = get_max_preds(y.detach().cpu().numpy())
    xy = xy.reshape(-1, 2)
    if div != 1:
        xy = xy / div                                                        # float32 divide, then truncate
    c = xy.astype(int)
    gt = lut[c[:, 0], c[:, 1], :, :].copy().reshape(b, k, oh, ow).copy()
    if variant == "base":
        off_diag = 1.0 - np.eye(k, dtype=np.float32)                         # regda_4.py:74
        gf = gt.reshape(b, k, -1).transpose((0, 2, 1))
        gf = gf.dot(off_diag).clip(max=1.0, min=0.0).transpose((0, 2, 1)).reshape(b, k, oh, ow).copy()
    else:
        gf = (np.ones_like(gt) - gt * 10).clip(max=1.0, min=0.0)             # regda_7.py:3036-3037
    return torch.from_numpy(gt).to(y.device), torch.from_numpy(gf).to(y.device)


# ----------------------------------------------------------------------------------------
# a8-a11  regression disparity                          regda_4.py:89-143 ; regda_7.py:3206-3268, 3485-3632
# ----------------------------------------------------------------------------------------

def _per_map_max_normalise(gf):
    """gf[b][k] / max(gf[b][k]) with the reference's B*K Python loop (regda_7.py:3546-3548,
    3623-3625); 0/0 stays NaN."""
    b, c = gf.shape[0], gf.shape[1]
    rows = [torch.stack([gf[s][j] / torch.max(gf[s][j]) for j in range(c)]) for s in range(b)]
    return torch.stack(rows)


def ground_maps(variant, y, y_adv2=None):
    """(gt, gf) exactly as the disparity variant builds them before calling its criterion."""
    if variant == "base":                                                    # regda_4.py:133-138
        return pseudo_label(y, "base")
    if variant == "x1":                                                      # regda_7.py:3250-3256
        gt, _ = pseudo_label(y, "01")
        gf = (torch.ones_like(gt) - gt * 10).clip(max=1.0, min=0.0)
        return gt, gf
    if variant == "x5":                                                      # regda_7.py:3529-3548
        gt, gf = pseudo_label(y, "03")
        gf1 = (torch.ones_like(gt) - gt * 10).clip(max=1.0, min=0.0)
        if y_adv2 is not None:
            gf = gf1 + y_adv2
            gf = (gf - gt * 100).clip(max=1.0, min=0.0)
        return gt, _per_map_max_normalise(gf)
    if variant == "x6":                                                      # regda_7.py:3609-3625
        gt, _ = pseudo_label(y, "base")
        lp = torch.sum(gt, dim=1).clip(max=1.0, min=0.0)
        lp = lp.unsqueeze(1).repeat(1, gt.shape[1], 1, 1)                    # reference hard-codes 21 (:3615)
        gf = (lp - gt * 10).clip(max=1.0, min=0.0)
        if y_adv2 is not None:
            gf = gf + y_adv2
            gf = (gf - gt * 100).clip(max=1.0, min=0.0)
        return gt, _per_map_max_normalise(gf)
    if variant == "rd4":                                                     # regda_4.py:344-356 (RegressionDisparity4)
        gt, _ = pseudo_label(y, "base")
        lp = torch.sum(gt, dim=1).clip(max=1.0, min=0.0)
        lp = lp.unsqueeze(1).repeat(1, gt.shape[1], 1, 1)
        return gt, (lp - gt * 10).clip(max=1.0, min=0.0)
    if variant in ("x2", "x3"):                                              # regda_7.py:3317-3337, 3385-3405
        gt, _ = pseudo_label(y, "base")                                      # (PseudoLabelGenerator02)
        return gt, (torch.ones_like(gt) - gt * 10).clip(max=1.0, min=0.0)
    if variant == "x4":                                                      # regda_7.py:3454-3482, y_adv2 = None
        gt, gf = pseudo_label(y, "01")
        return gt, _per_map_max_normalise(gf)
    raise ValueError(variant)


def _sample_max_normalise(lp):
    """``[lp[k] / max(lp[k]) for k in range(b)]`` stacked (regda_4.py:208-209)."""
    return torch.stack([lp[k] / torch.max(lp[k]) for k in range(lp.shape[0])])


def ground_maps_fusion(variant, y, label_1, label_2=None):
    """(gt, gf) of RegressionDisparity2/3/5/6/7/8 (regda_4.py:145-645): ``gf = clip(label_p - 10 gt)`` with the per-sample
    map ``label_p`` fused from the summed pseudo-labels of y, label_1 (and label_2)."""
    c01 = lambda t: t.clip(max=1.0, min=0.0)
    gt, _ = pseudo_label(y, "base")
    gt1, _ = pseudo_label(label_1, "base")
    gt2 = pseudo_label(label_2, "base")[0] if label_2 is not None else None
    if variant == "rd2":                                                     # :203-209
        lp = _sample_max_normalise(torch.sum(gt1, dim=1) + torch.sum(gt, dim=1) + torch.sum(gt2, dim=1))
    elif variant == "rd3":                                                   # :280-286
        lp = _sample_max_normalise(c01(torch.sum(gt, dim=1)) + c01(torch.sum(gt1, dim=1)) + c01(torch.sum(gt2, dim=1)))
    elif variant == "rd5":                                                   # :409-416
        p1, p2, p3 = c01(torch.sum(gt, dim=1)), c01(torch.sum(gt1, dim=1)), c01(torch.sum(gt2, dim=1))
        lp = c01(p1 + c01(p2 - p1) + c01(p3 - p1))
    elif variant == "rd6":                                                   # :479-484
        p1, p2 = c01(torch.sum(gt, dim=1)), c01(torch.sum(gt1, dim=1))
        lp = c01(p1 + c01(p2 - p1))
    elif variant == "rd7":                                                   # :553-558
        lp = _sample_max_normalise(c01(torch.sum(gt1, dim=1)) + c01(torch.sum(gt, dim=1)))
    elif variant == "rd8":                                                   # :625-634
        p1 = c01(torch.sum(gt, dim=1))
        x1 = c01(torch.sum(c01(gt1 - gt), dim=1))
        x2 = c01(torch.sum(c01(gt2 - gt), dim=1))
        lp = c01(p1 + x1 + x2)
    else:
        raise ValueError(variant)
    lp = lp.unsqueeze(1).repeat(1, gt.shape[1], 1, 1)                        # reference hard-codes 21
    return gt, (lp - gt * 10).clip(max=1.0, min=0.0)


def joints_mse_loss0(output, target, target_weight=None, reduction="mean"):
    """uda/model/loss.py:96-112: both maps shifted by 1e-7 and normalised to sum 1, then 0.5*(p-t)^2*w."""
    B, K = output.shape[0], output.shape[1]
    p = output.reshape((B, K, -1)) + 1e-7
    p = p / p.sum(dim=-1, keepdims=True)
    t = target.reshape((B, K, -1)) + 1e-7
    t = t / t.sum(dim=-1, keepdims=True)
    loss = F.mse_loss(p, t, reduction="none") * 0.5
    if target_weight is not None:
        loss = loss * target_weight.view((B, K, 1))
    if reduction == "mean":
        return loss.mean()
    if reduction == "none":
        return loss.mean(dim=-1)
    return None


def joints_kl_loss5(output, target, target_weight=None, reduction="mean", epsilon=0.0):
    """uda/model/loss.py:189-216: per-map rescale by the detached overlap score w5, then KL; target_weight is ignored."""
    B, K = output.shape[0], output.shape[1]
    f1 = (output / torch.max(output)).detach()
    f2 = (target / torch.max(target)).detach()
    w3 = torch.sum(torch.sum(torch.mul(f1, f2), dim=2), dim=2)
    w5 = (w3 / torch.max(w3)).unsqueeze(-1).unsqueeze(-1)
    output = torch.mul(output, w5)
    target = torch.mul(target, w5)
    logp = F.log_softmax(output.reshape((B, K, -1)), dim=-1)
    q = target.reshape((B, K, -1)) + epsilon
    q = q / q.sum(dim=-1, keepdims=True)
    loss = F.kl_div(logp, q, reduction="none").sum(dim=-1)
    if reduction == "mean":
        return loss.mean()
    if reduction == "none":
        return loss.mean(dim=-1)
    return None


def regression_disparity(variant, y, y_adv, y_adv2=None, weight=None, mode="min",
                         epsilon=1e-7, reduction="mean"):
    """criterion(y_adv, gt or gf, weight) with criterion = JointsKLLoss(epsilon) as wired at
    train1.py:135-137.  'min' -> ground truth, 'max' -> ground false."""
    assert mode in ("min", "max")
    gt, gf = ground_maps(variant, y.detach(), y_adv2)
    tgt = gt if mode == "min" else gf
    return joints_kl_loss(y_adv, tgt, weight, reduction=reduction, epsilon=epsilon)


# ----------------------------------------------------------------------------------------
# a12  multiscale fusion                                train1.py:410-424 (== test.py:362-376)
# ----------------------------------------------------------------------------------------

def fuse_multiscale(y_adv3, y_adv2, size_hi=64, size_mid=32):
    """target5 = 0.5*up_hi(y_adv3) + up_hi(y_adv2);  target0 = up_mid(y_adv3); bilinear,
    align_corners=False (nn.Upsample default).  train1.py:410-424."""
    up_hi3 = F.interpolate(y_adv3.detach(), size=size_hi, mode="bilinear")   # :410-411
    up_hi2 = F.interpolate(y_adv2.detach(), size=size_hi, mode="bilinear")   # :413-414
    up_mid3 = F.interpolate(y_adv3.detach(), size=size_mid, mode="bilinear") # :416-417
    return 0.5 * up_hi3 + up_hi2, up_mid3                                    # :424


def fuse_three_scales(lo, mid, hi):
    """Benchmark config 4 (BASELINE.json configs[3]): the same fusion rule scaled to 32/64/128:
    fused = 0.5*up(lo) + up(mid) + hi at the finest resolution."""
    size = hi.shape[-1]
    return (0.5 * F.interpolate(lo, size=size, mode="bilinear")
            + F.interpolate(mid, size=size, mode="bilinear") + hi)


# ----------------------------------------------------------------------------------------
# f4  decode helpers next to get_max_preds          utils/keypoint_detection.py:139-239
# ----------------------------------------------------------------------------------------

def find_keypoints_max(heatmaps: torch.Tensor) -> torch.Tensor:
    """[C,H,W] -> [C,3] (u, v, max).  utils/keypoint_detection.py:139-154: per-channel max over the flattened map;
    ``v = floor(ind / size(1))``, ``u = fmod(ind, size(2))`` in float32."""
    flat = heatmaps.reshape(heatmaps.size(0), -1)
    peak, where = flat.max(1)                                                # :148
    where = where.float()                                                    # :149
    v = torch.floor(torch.div(where, heatmaps.size(1)))                      # :151
    u = torch.fmod(where, heatmaps.size(2))                                  # :152
    return torch.cat((u.view(-1, 1), v.view(-1, 1), peak.view(-1, 1)), 1)    # :153


def _resize_bilinear(hm: torch.Tensor, resize_dim):
    """``nn.Upsample(size=resize_dim, mode='bilinear')`` == interpolate(..., align_corners=False)."""
    return F.interpolate(hm, size=resize_dim, mode="bilinear", align_corners=False)


def compute_uv_from_heatmaps(hm: torch.Tensor, resize_dim):
    """utils/keypoint_detection.py:156-171: resize, per-map argmax -> float [B,K,2] (u, v); no masking."""
    resized = _resize_bilinear(hm, resize_dim).view(-1, resize_dim[0], resize_dim[1])   # :163-165
    uvc = find_keypoints_max(resized).view(-1, hm.size(1), 3)                           # :168-169
    return uvc[:, :, :2]                                                                # :170


def compute_uv_from_heatmaps2(hm: torch.Tensor, resize_dim):
    """utils/keypoint_detection.py:174-205: resize, argmax / amax per map -> int64 [B,K,2] (x, y), zeroed where
    max <= 0 (the float quotient ``idx / width`` is truncated by the assignment into the int64 tensor)."""
    resized = _resize_bilinear(hm, resize_dim)
    n, k, w = resized.shape[0], resized.shape[1], resized.shape[3]
    flat = resized.reshape(n, k, -1)
    where = torch.argmax(flat, 2).reshape(n, k, 1)                           # :188
    peak = torch.amax(flat, 2).reshape(n, k, 1)                              # :189
    xy = where.repeat(1, 1, 2)                                               # :195
    xy[:, :, 0] = xy[:, :, 0] % w                                            # :197
    xy[:, :, 1] = xy[:, :, 1] / w                                            # :198 (true division, truncated on store)
    xy *= torch.greater(peak, 0.0).repeat(1, 1, 2)                           # :201-205
    return xy


def compute_uv_from_heatmaps3(heatmap: torch.Tensor) -> torch.Tensor:
    """utils/keypoint_detection.py:209-239: softmax(100*h) over the map; expectation of the row index (``xx``) and the
    column index (``yy``); output ``cat([E[col], E[row]]) * 4``."""
    scaled = heatmap.mul(100)                                                # :214
    n, k, h, w = scaled.size()
    p = F.softmax(scaled.view(n, k, h * w), dim=2).view(n, k, h, w)          # :218-220
    rows, cols = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")   # :222 (xx = row, yy = col)
    e_row = p.mul(rows.float().to(heatmap.device)).view(n, k, h * w).sum(2).unsqueeze(2)   # :224-229
    e_col = p.mul(cols.float().to(heatmap.device)).view(n, k, h * w).sum(2).unsqueeze(2)   # :230-235
    return torch.cat([e_col, e_row], 2) * 4                                  # :237-239


# ----------------------------------------------------------------------------------------
# The benchmarked pipeline: gen + loss + decode + PCK   (BASELINE.json configs[0], configs[1])
# ----------------------------------------------------------------------------------------

def pipeline(pred: np.ndarray, joints: np.ndarray, joints_vis: np.ndarray, sigma=2,
             image_size=(256, 256), kl_epsilon=0.0, thr=0.5):
    """generate_target x B  ->  JointsMSELoss + JointsKLLoss  ->  accuracy (2x decode + PCK).

    pred float32[B,K,H,W]; joints float64[B,K,2] image px; joints_vis [B,K,1].
    Returns a dict with the scalars / vectors the fused CUDA pipeline must reproduce."""
    b, k, h, w = pred.shape
    target, weight = generate_target_batch(joints, joints_vis, (w, h), sigma, image_size)
    tp, tt, tw = torch.from_numpy(pred), torch.from_numpy(target), torch.from_numpy(weight)
    mse = joints_mse_loss(tp, tt, tw)
    kl = joints_kl_loss(tp, tt, tw, epsilon=kl_epsilon)
    acc, avg, cnt, xy = accuracy(pred, target, thr=thr)
    t_xy, _ = get_max_preds(target)
    hits, valid = pck_counts(xy, t_xy, h, w, thr)
    return dict(mse=float(mse), kl=float(kl), acc=acc, avg_acc=float(avg), cnt=int(cnt),
                pred_xy=xy, hits=hits, valid=valid, target=target, weight=weight)
