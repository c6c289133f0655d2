"""Seeded parity cases for every row of SURVEY.md §8(a)  (TEST INFRASTRUCTURE).

Each case takes an implementation namespace ``ns`` carrying the reference's call surface
(the real reference via ``ref_loader.load()``, the oracle via ``oracle.api.namespace()`` or the
CUDA package) plus a torch ``device`` for the tensor arguments, and returns
``{key: ndarray}``.  Key prefix states the parity bar (SURVEY.md §8c tolerances):

* ``x:``  bit-exact   (indices, coordinates, maxvals, PCK counts, cnt, weights)
* ``c:``  close       (fp32 heatmaps / losses / gradients: rtol 1e-5, atol 1e-6)

``oracle/gen_golden.py`` evaluates the cases on the REAL reference and freezes the outputs
under ``tests/golden/``; inputs are never stored, they are regenerated from the seed.
"""
from __future__ import annotations

import importlib
import os
import sys

import numpy as np
import torch

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
synth = importlib.import_module("domain-adaptative-hand-pose-estimation_b200.synth")

RTOL, ATOL = 1e-5, 1e-6
K = 21


def _np(t):
    if isinstance(t, torch.Tensor):
        return t.detach().cpu().numpy()
    return np.asarray(t)


def _t(a, device, grad=False):
    t = torch.from_numpy(np.ascontiguousarray(a)).to(device)
    if grad:
        t.requires_grad_(True)
    return t


# ---------------------------------------------------------------------------------- a1

def special_maps(H=64, W=64):
    """Hand-made decode edge cases (SURVEY.md §8c known answers)."""
    m = []
    a = np.zeros((H, W), np.float32); a[5, 9] = 0.7; m.append(a)                  # single peak -> (9,5)
    m.append(np.full((H, W), -1.0, np.float32))                                     # all negative -> (0,0)
    m.append(np.zeros((H, W), np.float32))                                          # all zero -> (0,0), max 0
    a = np.zeros((H, W), np.float32); a[3, 4] = 2.0; a[40, 1] = 2.0; m.append(a)    # duplicate max -> first
    a = np.full((H, W), 0.25, np.float32); m.append(a)                              # constant -> idx 0
    a = np.zeros((H, W), np.float32); a[H - 1, W - 1] = 1e-30; m.append(a)          # tiny positive at the end
    a = np.zeros((H, W), np.float32); a[7, 7] = 3.0; a[20, 30] = np.nan; a[50, 2] = np.nan
    m.append(a)                                                                     # NaN -> first NaN wins, masked
    a = np.zeros((H, W), np.float32); a[0, 0] = -0.0; a[1, 1] = 0.0; m.append(a)    # signed zeros
    a = np.full((H, W), -np.inf, np.float32); a[10, 11] = -5.0; m.append(a)         # -inf background
    a = np.zeros((H, W), np.float32); a[2, 3] = np.inf; a[2, 2] = 1e38; m.append(a) # +inf
    a = np.zeros((H, W), np.float32); a[0, 1] = 1.0; m.append(a)                    # peak at x=1,y=0 (PCK invalid)
    a = np.zeros((H, W), np.float32); a[H - 1, 0] = 1.0; m.append(a)                # bottom-left corner
    return np.stack(m)


def case_decode(ns, device="cpu"):
    out = {}
    d = synth.make_host_batch(101, 3, K, 64, 64)
    p, mv = ns.get_max_preds(d["pred"])
    out["x:preds64"], out["x:maxvals64"] = _np(p), _np(mv)
    sp = special_maps()
    n = sp.shape[0]
    reps = -(-K // n)
    heat = np.concatenate([sp] * reps)[:K][None]                 # [1,K,64,64]
    heat = np.concatenate([heat, heat[:, ::-1]], axis=0).copy()  # second sample reversed joint order
    p, mv = ns.get_max_preds(heat)
    out["x:preds_special"], out["x:maxvals_special"] = _np(p), _np(mv)
    for s, seed in ((16, 102), (32, 103), (128, 104)):
        d = synth.make_host_batch(seed, 2, K, s, s)
        p, mv = ns.get_max_preds(d["pred"])
        out[f"x:preds{s}"], out[f"x:maxvals{s}"] = _np(p), _np(mv)
    d = synth.make_host_batch(105, 2, 5, 48, 80)                 # non-square, K != 21
    p, mv = ns.get_max_preds(d["pred"])
    out["x:preds_48x80"], out["x:maxvals_48x80"] = _np(p), _np(mv)
    return out


# ---------------------------------------------------------------------------------- a5

def target_inputs(seed=201, B=4, image_size=256):
    rs = np.random.RandomState(seed)
    joints = rs.uniform(-8.0, image_size + 8.0, size=(B, K, 2))
    vis = (rs.uniform(size=(B, K, 1)) < 0.9).astype(np.float32)
    # hand-placed edge cases in sample 0
    edge = [(0.0, 0.0), (255.9, 255.9), (-1.9, 10.0), (-6.1, 10.0), (257.9, 100.0), (258.0, 100.0),
            (2.0, 2.0), (1.99, 6.0), (128.0, 0.0), (0.0, 128.0), (253.0, 3.0), (24.0, 250.0),
            (-5.99, -5.99), (100.5, 100.49)]
    for i, (x, y) in enumerate(edge):
        joints[0, i] = (x, y)
        vis[0, i, 0] = 1.0
    vis[0, 14, 0] = 0.5      # exactly 0.5 -> not pasted (util.py:64 is a strict >)
    vis[0, 15, 0] = 0.51     # fractional visibility -> pasted, weight stays 0.51
    vis[0, 16, 0] = 0.0
    return joints, vis


def case_target(ns, device="cpu"):
    out = {}
    joints, vis = target_inputs()
    for s in (64, 32, 128):
        ts, ws = [], []
        for b in range(joints.shape[0]):
            t, w = ns.generate_target(joints[b], vis[b], (s, s), 2, (256, 256))
            ts.append(_np(t)); ws.append(_np(w))
        out[f"c:target{s}"], out[f"x:weight{s}"] = np.stack(ts), np.stack(ws)
    t, w = ns.generate_target(joints[1][:7], vis[1][:7], (48, 64), 1, (192, 256))  # W=48,H=64, sigma 1
    out["c:target_48x64_s1"], out["x:weight_48x64_s1"] = _np(t), _np(w)
    t, w = ns.generate_target(joints[2], vis[2], (64, 64), 3, (256, 256))          # sigma 3 (19x19 patch)
    out["c:target64_s3"], out["x:weight64_s3"] = _np(t), _np(w)
    return out


# ---------------------------------------------------------------------------------- a2

def _targets_for(ns_oracle_target, d, s):
    ts, ws = [], []
    for b in range(d["joints"].shape[0]):
        t, w = ns_oracle_target(d["joints"][b], d["vis"][b], (s, s), 2, (256, 256))
        ts.append(_np(t)); ws.append(_np(w))
    return np.stack(ts).astype(np.float32), np.stack(ws).astype(np.float32)


def _oracle_target():
    from . import hp_oracle
    return hp_oracle.generate_target


def case_accuracy(ns, device="cpu"):
    """Targets always come from the oracle's generate_target so this case isolates a1+a2."""
    out = {}
    for s, seed, B in ((64, 301, 6), (32, 302, 4), (128, 303, 2), (16, 304, 4)):
        d = synth.make_host_batch(seed, B, K, s, s)
        tgt, _ = _targets_for(_oracle_target(), d, s)
        acc, avg, cnt, pred = ns.accuracy(d["pred"], tgt)
        out[f"x:acc{s}"] = np.asarray(acc, dtype=np.float64)
        out[f"x:avg{s}"] = np.asarray([avg], dtype=np.float64)
        out[f"x:cnt{s}"] = np.asarray([cnt], dtype=np.int64)
        out[f"x:pred{s}"] = _np(pred)
    # no valid joint at all -> acc = -1 everywhere, avg 0, cnt 0
    z = np.zeros((2, K, 64, 64), np.float32)
    acc, avg, cnt, pred = ns.accuracy(z, z)
    out["x:acc_empty"] = np.asarray(acc, dtype=np.float64)
    out["x:avg_empty"] = np.asarray([avg], dtype=np.float64)
    out["x:cnt_empty"] = np.asarray([cnt], dtype=np.int64)
    # threshold edge at 64x64: dx^2+dy^2 = 10 hits, 11 does not
    o = np.zeros((1, K, 64, 64), np.float32); t = np.zeros((1, K, 64, 64), np.float32)
    offs = [(3, 1), (1, 3), (3, 2), (0, 3), (2, 2), (0, 4), (-3, -1), (-1, 3)]
    for j in range(K):
        dx, dy = offs[j % len(offs)]
        t[0, j, 30, 30] = 1.0
        o[0, j, 30 + dy, 30 + dx] = 1.0
    acc, avg, cnt, pred = ns.accuracy(o, t)
    out["x:acc_edge"] = np.asarray(acc, dtype=np.float64)
    out["x:avg_edge"] = np.asarray([avg], dtype=np.float64)
    return out


# ---------------------------------------------------------------------------------- a3 / a4

def loss_inputs(seed, B, s, device):
    d = synth.make_host_batch(seed, B, K, s, s)
    tgt, w = _targets_for(_oracle_target(), d, s)
    return d["pred"], tgt, w


def case_losses(ns, device="cpu"):
    out = {}
    for s, seed, B in ((64, 401, 4), (32, 402, 3), (16, 403, 3), (128, 404, 1)):
        p, t, w = loss_inputs(seed, B, s, device)
        for name, make in (("mse", lambda red: ns.JointsMSELoss(reduction=red)),
                           ("kl0", lambda red: ns.JointsKLLoss(reduction=red, epsilon=0.0)),
                           ("kl7", lambda red: ns.JointsKLLoss(reduction=red, epsilon=1e-7))):
            for red in ("mean", "none"):
                for wname, wt in (("w", w), ("nw", None)):
                    if name == "kl0" and s != 64:
                        continue
                    tp = _t(p, device, grad=True)
                    tt = _t(t, device)
                    tw = None if wt is None else _t(wt, device)
                    crit = make(red)
                    if hasattr(crit, "to"):
                        crit = crit.to(device)
                    l = crit(tp, tt, tw)
                    out[f"c:{name}_{red}_{wname}_{s}"] = _np(l)
                    if (s == 16 or (s == 32 and red == "mean" and wname == "w")) and name != "kl0":
                        rs = np.random.RandomState(seed + 7)
                        go = np.asarray(rs.uniform(0.5, 1.5, size=tuple(l.shape)), dtype=np.float32)
                        go = torch.from_numpy(go.copy()).reshape(tuple(l.shape)).to(device)
                        l.backward(go)
                        out[f"c:grad_{name}_{red}_{wname}_{s}"] = _np(tp.grad)
    # dense (non-Gaussian) positive target, the shape 'max'-mode ground-false maps have
    rs = np.random.RandomState(405)
    p = rs.standard_normal((2, K, 64, 64)).astype(np.float32)
    t = rs.uniform(0, 1, size=(2, K, 64, 64)).astype(np.float32)
    w = rs.uniform(0, 1, size=(2, K, 1)).astype(np.float32)
    out["c:kl7_dense_mean"] = _np(ns.JointsKLLoss(epsilon=1e-7)(_t(p, device), _t(t, device), _t(w, device)))
    out["c:mse_dense_mean"] = _np(ns.JointsMSELoss()(_t(p, device), _t(t, device), _t(w, device)))
    return out


# ---------------------------------------------------------------------------------- a6 / a7

def case_pseudo_label(ns, device="cpu"):
    out = {}
    d = synth.make_host_batch(501, 2, K, 64, 64)
    y = _t(d["pred"], device)
    for name, cls in (("base", ns.PseudoLabelGenerator), ("03", ns.PseudoLabelGenerator03),
                      ("01", ns.PseudoLabelGenerator01)):
        if name == "base":
            plg = cls(K, 64, 64)
        else:
            plg = cls(K)
        gt, gf = plg(y)
        out[f"c:gt_{name}"], out[f"c:gf_{name}"] = _np(gt), _np(gf)
    d32 = synth.make_host_batch(502, 2, K, 32, 32)
    gt, gf = ns.PseudoLabelGenerator(K, 32, 32)(_t(d32["pred"], device))           # base at another size
    out["c:gt_base32"], out["c:gf_base32"] = _np(gt), _np(gf)
    return out


# ---------------------------------------------------------------------------------- a8 - a11

def disparity_inputs(seed=601, B=2):
    d = synth.make_host_batch(seed, B, K, 64, 64)
    adv = synth.make_host_batch(seed + 1, B, K, 64, 64)["pred"]
    adv32, adv16 = synth.make_lowres_heads(seed + 2, adv, (32, 16))
    rs = np.random.RandomState(seed + 3)
    w = (rs.uniform(size=(B, K, 1)) < 0.85).astype(np.float32)
    fused64 = np.clip(synth.make_host_batch(seed + 4, B, K, 64, 64)["pred"], -0.2, 1.2).astype(np.float32)
    fused32 = np.clip(synth.make_host_batch(seed + 5, B, K, 32, 32)["pred"], -0.2, 1.2).astype(np.float32)
    return dict(y=d["pred"], adv64=adv, adv32=adv32, adv16=adv16, w=w, fused64=fused64, fused32=fused32)


def case_disparity(ns, device="cpu"):
    out = {}
    I = disparity_inputs()
    y = _t(I["y"], device)
    w = _t(I["w"], device)
    kl = lambda: ns.JointsKLLoss(epsilon=1e-7)
    variants = (
        ("base", lambda: ns.RegressionDisparity(ns.PseudoLabelGenerator(K, 64, 64), kl()), "adv64", None),
        ("x1", lambda: ns.RegressionDisparityx1(ns.PseudoLabelGenerator01(K), kl()), "adv16", None),
        ("x5", lambda: ns.RegressionDisparityx5(ns.PseudoLabelGenerator03(K), kl()), "adv32", "fused32"),
        ("x6", lambda: ns.RegressionDisparityx6(ns.PseudoLabelGenerator(K, 64, 64), kl()), "adv64", "fused64"),
    )
    for name, make, adv_key, fused_key in variants:
        rd = make()
        for mode in ("min", "max"):
            for fz in ((None,) if fused_key is None else (None, fused_key)):
                for wname, wt in (("w", w), ("nw", None)):
                    adv = _t(I[adv_key], device, grad=True)
                    if name in ("base", "x1"):
                        l = rd(y, adv, wt, mode)
                    else:
                        f = None if fz is None else _t(I[fz], device)
                        l = rd(y, adv, f, wt, mode)
                    tag = f"{name}_{mode}_{'f' if fz else 'nf'}_{wname}"
                    out[f"c:loss_{tag}"] = _np(l)
                    if wname == "w":
                        l.backward()
                        out[f"c:grad_{tag}"] = _np(adv.grad)
                        out[f"c:gt_{tag}"] = _np(rd.ground_truth)
                        out[f"c:gf_{tag}"] = _np(rd.ground_false)
    return out


# ---------------------------------------------------------------------------------- a12

def fusion_inputs(seed=701, B=2, Kf=5):
    rs = np.random.RandomState(seed)
    y3 = rs.standard_normal((B, Kf, 16, 16)).astype(np.float32)
    y2 = rs.standard_normal((B, Kf, 32, 32)).astype(np.float32)
    return y3, y2


def reference_fusion(y3, y2):
    """The inline statements of train1.py:410-424, with torch's own modules."""
    import torch.nn as nn
    target = nn.Upsample(size=64, mode="bilinear")(y3.detach())
    target1 = nn.Upsample(size=64, mode="bilinear")(y2.detach())
    target0 = nn.Upsample(size=32, mode="bilinear")(y3.detach())
    return 0.5 * target + target1, target0


def case_fusion(ns, device="cpu"):
    y3, y2 = fusion_inputs()
    fuse = getattr(ns, "fuse_multiscale", None) or reference_fusion
    t5, t0 = fuse(_t(y3, device), _t(y2, device))
    return {"c:target5": _np(t5), "c:target0": _np(t0)}


# ---------------------------------------------------------------------------------- step B of train(), as written

def case_step_b(ns, device="cpu"):
    """train1.py:405-431 (== test.py:357-383): the three ground-false losses of step B from the adversarial heads - x1 'max' on the
    16x16 head, x6 'max' against ``target5 = 0.5 * up64(y_adv3) + up64(y_adv2)``, x5 'max' against ``target0 = up32(y_adv3)`` -,
    their weighted sum (train1.py:430) and its gradients with respect to the three heads.  The reference and the oracle execute
    the driver's inline ``nn.Upsample`` statements; a namespace that offers ``FusedHeads`` (the CUDA package) hands ``target5``
    to the loss UNFUSED, so the golden values of this case pin the fusion done inside the loss kernel to the real reference."""
    I = disparity_inputs(seed=801, B=3)
    y, w = _t(I["y"], device), _t(I["w"], device)
    y_adv, y_adv2, y_adv3 = (_t(I[k], device, grad=True) for k in ("adv64", "adv32", "adv16"))
    kl = lambda: ns.JointsKLLoss(epsilon=1e-7)                                   # train1.py:135-137
    regression_disparity = ns.RegressionDisparityx6(ns.PseudoLabelGenerator(K, 64, 64), kl())
    regression_disparity2 = ns.RegressionDisparityx5(ns.PseudoLabelGenerator03(K), kl())
    regression_disparity1 = ns.RegressionDisparityx1(ns.PseudoLabelGenerator01(K), kl())
    loss1 = regression_disparity1(y, y_adv3, w, mode="max")                      # train1.py:408
    if getattr(ns, "FusedHeads", None) is not None:
        target5 = ns.FusedHeads(y_adv3, y_adv2)
        target0 = ns.upsample_bilinear(y_adv3.detach(), 32)
    else:
        target5, target0 = reference_fusion(y_adv3, y_adv2)                      # train1.py:410-424
    loss2 = regression_disparity(y, y_adv, target5, w, mode="max")               # train1.py:426
    loss3 = regression_disparity2(y, y_adv2, target0, w, mode="max")             # train1.py:428
    total = 0.3 * loss1 + 1 * loss2 + 0.3 * loss3                                # train1.py:430
    total.backward()
    return {"c:loss1": _np(loss1), "c:loss2": _np(loss2), "c:loss3": _np(loss3), "c:total": _np(total),
            "c:grad_y_adv": _np(y_adv.grad), "c:grad_y_adv2": _np(y_adv2.grad), "c:grad_y_adv3": _np(y_adv3.grad)}


# ---------------------------------------------------------------------------------- f4

def case_soft_decode(ns, device="cpu"):
    """Decode helpers of utils/keypoint_detection.py:139-239 (row f4).  The resize of (1)/(2) is ATen's on the
    reference side and the fusion kernel's on the CUDA side (<= 4.8e-7 apart), so the argmax
    outputs are bit-exact only while the top-2 margin of a resized map exceeds that; the seeded maps below have
    clear peaks."""
    out = {}
    d = synth.make_host_batch(1201, 2, K, 64, 64)
    hm = d["pred"]
    hm[0, 0] = -np.abs(hm[0, 0]) - 0.1                                              # all negative: (2) masks it
    lo = synth.make_lowres_heads(1202, hm, (32, 16))
    t64, t32 = _t(hm, device), _t(lo[0], device)
    out["x:uv1_64to128"] = _np(ns.compute_uv_from_heatmaps(t64, (128, 128)))
    out["x:uv1_32to64"] = _np(ns.compute_uv_from_heatmaps(t32, (64, 64)))
    out["x:uv2_32to64"] = _np(ns.compute_uv_from_heatmaps2(t32, (64, 64)))
    out["x:uv2_64to128"] = _np(ns.compute_uv_from_heatmaps2(t64, (128, 128)))
    fk = _np(ns.find_keypoints_max(_t(hm[1], device)))
    out["x:fkm_uv"], out["x:fkm_max"] = fk[:, :2], fk[:, 2]
    out["c:uv3_64"] = _np(ns.compute_uv_from_heatmaps3(t64))
    out["c:uv3_32"] = _np(ns.compute_uv_from_heatmaps3(t32))
    rs = np.random.RandomState(1203)
    flat = (0.02 * rs.standard_normal((2, 5, 30, 42))).astype(np.float32)            # broad soft-argmax, odd size
    out["c:uv3_30x42"] = _np(ns.compute_uv_from_heatmaps3(_t(flat, device)))
    return out


# ---------------------------------------------------------------------------------- f3

def case_variants(ns, device="cpu"):
    """the disparity variants the drivers import but never call (RegressionDisparity2-8, x2-x4)
    and the loss weightings JointsMSELoss0 / JointsKLLoss5, every one with loss, gradient and both maps."""
    out = {}
    I = disparity_inputs(seed=1301)
    y, w = _t(I["y"], device), _t(I["w"], device)
    l1 = _t(synth.make_host_batch(1311, 2, K, 64, 64)["pred"], device)
    l2 = _t(synth.make_host_batch(1312, 2, K, 64, 64)["pred"], device)
    kl = lambda: ns.JointsKLLoss(epsilon=1e-7)
    plg = lambda: ns.PseudoLabelGenerator(K, 64, 64)
    specs = (
        ("rd2", lambda: ns.RegressionDisparity2(plg(), kl()), "adv64", (l1, l2)),
        ("rd3", lambda: ns.RegressionDisparity3(plg(), kl()), "adv64", (l1, l2)),
        ("rd4", lambda: ns.RegressionDisparity4(plg(), kl()), "adv64", ()),
        ("rd5", lambda: ns.RegressionDisparity5(plg(), kl()), "adv64", (l1, l2)),
        ("rd6", lambda: ns.RegressionDisparity6(plg(), kl()), "adv64", (l1,)),
        ("rd7", lambda: ns.RegressionDisparity7(plg(), kl()), "adv64", (l1,)),
        ("rd8", lambda: ns.RegressionDisparity8(plg(), kl()), "adv64", (l1, l2)),
        ("x2", lambda: ns.RegressionDisparityx2(ns.PseudoLabelGenerator02(K, 64, 64), kl()), "adv64", ()),
        ("x3", lambda: ns.RegressionDisparityx3(ns.PseudoLabelGenerator02(K, 64, 64), kl()), "adv64", ()),
        ("x4", lambda: ns.RegressionDisparityx4(ns.PseudoLabelGenerator01(K), kl()), "adv16", ()),
    )
    for name, make, adv_key, labels in specs:
        rd = make()
        for mode in ("min", "max"):
            adv = _t(I[adv_key], device, grad=True)
            l = rd(y, adv, *labels, w, mode=mode) if name != "x4" else rd(y, adv, w, None, mode)
            out[f"c:loss_{name}_{mode}"] = _np(l)
            l.backward()
            out[f"c:grad_{name}_{mode}"] = _np(adv.grad)
            if mode == "max":
                out[f"c:gt_{name}"] = _np(rd.ground_truth)
                out[f"c:gf_{name}"] = _np(rd.ground_false)
    # loss weightings on positive maps (JointsMSELoss0 normalises by the map sum)
    rs = np.random.RandomState(1320)
    p = rs.uniform(0.01, 1.0, size=(2, K, 64, 64)).astype(np.float32)
    t = np.clip(synth.make_host_batch(1321, 2, K, 64, 64)["pred"], 0.0, None).astype(np.float32) + np.float32(1e-3)
    wt = rs.uniform(0, 1, size=(2, K, 1)).astype(np.float32)
    for red in ("mean", "none"):
        for wname, wv in (("w", wt), ("nw", None)):
            tp = _t(p, device, grad=True)
            l = ns.JointsMSELoss0(reduction=red)(tp, _t(t, device), None if wv is None else _t(wv, device))
            out[f"c:mse0_{red}_{wname}"] = _np(l)
            l.sum().backward()
            out[f"c:grad_mse0_{red}_{wname}"] = _np(tp.grad)
            tp = _t(p, device, grad=True)
            l = ns.JointsKLLoss5(reduction=red, epsilon=1e-7)(tp, _t(t, device), None if wv is None else _t(wv, device))
            out[f"c:kl5_{red}_{wname}"] = _np(l)
            l.sum().backward()
            out[f"c:grad_kl5_{red}_{wname}"] = _np(tp.grad)
    return out


CASES = {
    "decode": case_decode,
    "target": case_target,
    "accuracy": case_accuracy,
    "losses": case_losses,
    "pseudo_label": case_pseudo_label,
    "disparity": case_disparity,
    "fusion": case_fusion,
    "step_b": case_step_b,
    "soft_decode": case_soft_decode,
    "variants": case_variants,
}


def compare(got: dict, want: dict, rtol=RTOL, atol=ATOL):
    """Raise AssertionError naming the first key that misses its parity bar."""
    missing = sorted(set(want) - set(got))
    assert not missing, f"missing outputs: {missing}"
    for key in sorted(want):
        g, w = np.asarray(got[key]), np.asarray(want[key])
        assert g.shape == w.shape, f"{key}: shape {g.shape} != {w.shape}"
        if key.startswith("x:"):
            same = (g == w) | (np.isnan(g.astype(np.float64)) & np.isnan(w.astype(np.float64)))
            assert same.all(), f"{key}: {int((~same).sum())} of {same.size} entries differ (bit-exact bar)"
        else:
            a = atol
            if "grad" in key and w.size:
                # gradients are c*(softmax - q): both terms carry ~1e-7 relative fp32 error (the reference's
                # too) and nearly cancel, so the bar is 1e-5 of the tensor's scale, not of each element
                finite = np.abs(w[np.isfinite(w)])
                a = 1e-5 * float(finite.max()) if finite.size else atol
            np.testing.assert_allclose(g, w, rtol=rtol, atol=a, equal_nan=True, err_msg=key)
