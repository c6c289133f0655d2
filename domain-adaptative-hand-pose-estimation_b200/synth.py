"""Seeded synthetic heatmap workloads - no dataset or checkpoint is reachable
offline, so tests, golden fixtures and ``bench.py`` all draw from this recipe.

* keypoints ``U(-8, image+8)`` image px  (about 6 % land out of bounds -> weight-0 path of
  ``generate_target``, uda/dataset/util.py:42-46), visibility ``Bernoulli(0.9)``;
* prediction ``a * G(mu + round(N(0, 2.5^2))) + 0.05 * N(0, 1)`` with ``a ~ U(0.6, 1)`` so PCK is
  neither 0 nor 1; 1 % of the maps are forced all <= 0 (decode mask path,
  utils/keypoint_detection.py:31-34) and 0.5 % carry an exactly duplicated maximum (first-index
  tie-break of numpy argmax).

:func:`make_host_batch` is numpy ``RandomState`` (a frozen stream, so committed goldens can be
regenerated from the seed alone); :func:`make_device_batch` is the same recipe on the GPU for
bench-sized inputs.
"""
from __future__ import annotations

import numpy as np


def _centres(joints, w, h, image_size):
    stride_x = image_size / w
    stride_y = image_size / h
    mu = np.empty(joints.shape, dtype=np.int64)
    mu[..., 0] = np.trunc(joints[..., 0] / stride_x + 0.5)
    mu[..., 1] = np.trunc(joints[..., 1] / stride_y + 0.5)
    return mu


def make_host_batch(seed, B, K=21, H=64, W=64, image_size=256, sigma=2.0,
                    frac_nonpositive=0.01, frac_dup_max=0.005):
    """-> dict(pred f32[B,K,H,W], joints f64[B,K,2], vis f32[B,K,1])."""
    rs = np.random.RandomState(seed)
    joints = rs.uniform(-8.0, image_size + 8.0, size=(B, K, 2))
    vis = (rs.uniform(size=(B, K, 1)) < 0.9).astype(np.float32)
    mu = _centres(joints, W, H, image_size)
    jitter = np.rint(rs.normal(0.0, 2.5, size=(B, K, 2))).astype(np.int64)
    c = mu + jitter
    cx = np.clip(c[..., 0], 0, W - 1).astype(np.float32)[..., None, None]
    cy = np.clip(c[..., 1], 0, H - 1).astype(np.float32)[..., None, None]
    amp = rs.uniform(0.6, 1.0, size=(B, K, 1, 1)).astype(np.float32)
    xs = np.arange(W, dtype=np.float32)[None, None, None, :]
    ys = np.arange(H, dtype=np.float32)[None, None, :, None]
    pred = amp * np.exp(-((xs - cx) ** 2 + (ys - cy) ** 2) / np.float32(2.0 * sigma * sigma))
    pred = (pred + np.float32(0.05) * rs.standard_normal((B, K, H, W)).astype(np.float32)).astype(np.float32)
    sel = rs.uniform(size=(B, K))
    neg = sel < frac_nonpositive
    pred[neg] = -np.abs(pred[neg])
    dup = (sel >= frac_nonpositive) & (sel < frac_nonpositive + frac_dup_max)
    flat = pred.reshape(B, K, -1)
    for b, k in zip(*np.nonzero(dup)):
        i, j = sorted(rs.choice(H * W, size=2, replace=False))
        top = flat[b, k].max() + np.float32(1.0)
        flat[b, k, i] = top
        flat[b, k, j] = top
    return dict(pred=pred, joints=joints, vis=vis)


def make_lowres_heads(seed, pred, sizes=(32, 16)):
    """Adversarial-head style low-resolution maps: average-pooled prediction + noise."""
    rs = np.random.RandomState(seed)
    B, K, H, W = pred.shape
    out = []
    for s in sizes:
        f = H // s
        pooled = pred.reshape(B, K, s, f, s, f).mean(axis=(3, 5))
        out.append((pooled + np.float32(0.05) * rs.standard_normal(pooled.shape).astype(np.float32))
                   .astype(np.float32))
    return out


def make_device_batch(seed, B, K=21, H=64, W=64, image_size=256, sigma=2.0, device="cuda",
                      frac_nonpositive=0.01, frac_dup_max=0.005):
    """Same recipe with a ``torch.Generator`` on ``device`` (setup code, never timed).
    -> dict(pred f32[B,K,H,W], joints f64[B,K,2], vis f32[B,K,1]) resident on ``device``."""
    import torch

    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    joints = torch.rand((B, K, 2), generator=g, device=device, dtype=torch.float64) * (image_size + 16.0) - 8.0
    vis = (torch.rand((B, K, 1), generator=g, device=device) < 0.9).to(torch.float32)
    stride = torch.tensor([image_size / W, image_size / H], device=device, dtype=torch.float64)
    mu = torch.trunc(joints / stride + 0.5)
    jitter = torch.round(torch.randn((B, K, 2), generator=g, device=device, dtype=torch.float64) * 2.5)
    c = mu + jitter
    cx = c[..., 0].clamp(0, W - 1).to(torch.float32)[..., None, None]
    cy = c[..., 1].clamp(0, H - 1).to(torch.float32)[..., None, None]
    amp = torch.rand((B, K, 1, 1), generator=g, device=device) * 0.4 + 0.6
    xs = torch.arange(W, device=device, dtype=torch.float32)[None, None, None, :]
    ys = torch.arange(H, device=device, dtype=torch.float32)[None, None, :, None]
    pred = torch.empty((B, K, H, W), device=device, dtype=torch.float32)
    step = max(1, (1 << 26) // (K * H * W))          # build in slabs: bounded temporaries
    for b0 in range(0, B, step):
        b1 = min(B, b0 + step)
        gauss = amp[b0:b1] * torch.exp(-((xs - cx[b0:b1]) ** 2 + (ys - cy[b0:b1]) ** 2) / (2.0 * sigma * sigma))
        noise = torch.randn((b1 - b0, K, H, W), generator=g, device=device)
        pred[b0:b1] = gauss + 0.05 * noise
    sel = torch.rand((B, K), generator=g, device=device)
    neg = sel < frac_nonpositive
    pred[neg] = -pred[neg].abs()
    dup = ((sel >= frac_nonpositive) & (sel < frac_nonpositive + frac_dup_max)).nonzero()
    flat = pred.view(B, K, -1)
    if dup.numel():
        pos = torch.randint(0, H * W, (dup.shape[0], 2), generator=g, device=device)
        top = flat[dup[:, 0], dup[:, 1]].max(dim=-1).values + 1.0
        flat[dup[:, 0], dup[:, 1], pos[:, 0]] = top
        flat[dup[:, 0], dup[:, 1], pos[:, 1]] = top
    return dict(pred=pred, joints=joints, vis=vis)
