"""Gaussian target generation on the GPU (``uda/dataset/util.py:9-68``).

``generate_target`` keeps the reference's per-sample numpy signature (the datasets call it from
``__getitem__``: ``hand_3d_studio.py:102``, ``STB.py:150``, ``rendered_hand_pose.py:83``);
``generate_target_batch`` is the B200-native form: keypoints of a whole batch in, targets resident
on the device out (no per-sample numpy, no host->device copy of labels)."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError("no CUDA device: the B200 heatmap path has no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def generate_target_batch(joints, joints_vis, heatmap_size, sigma, image_size, device=None):
    """joints [B,K,2] (image px, float64), joints_vis [B,K,1] or [B,K]; heatmap_size (W,H);
    -> (target float32 [B,K,H,W], target_weight float32 [B,K,1]) CUDA tensors."""
    device = torch.device(device) if device is not None else (
        joints.device if isinstance(joints, torch.Tensor) and joints.is_cuda else _device())
    j = torch.as_tensor(joints).to(device=device, dtype=torch.float64).contiguous()
    v = torch.as_tensor(joints_vis).to(device=device, dtype=torch.float32).contiguous()
    if j.ndim != 3 or j.shape[-1] != 2:
        raise ValueError("joints must be [B,K,2]")
    B, K = j.shape[0], j.shape[1]
    if v.numel() != B * K:
        raise ValueError("joints_vis must have B*K elements")
    W, H = int(heatmap_size[0]), int(heatmap_size[1])
    tmp = _lib.integer_tmp(sigma * 3)                                   # util.py:30
    stride = np.array(image_size) / np.array(heatmap_size)              # util.py:36 (float64)
    target = torch.empty((B, K, H, W), dtype=torch.float32, device=device)
    weight = torch.empty((B, K, 1), dtype=torch.float32, device=device)
    with _lib.on_device(device):
        tab = _lib.gaussian_table(sigma, tmp, device)
        _lib.call("hp_gaussian_target", _lib.ptr(j), _lib.ptr(v), B * K, H, W, C.c_double(float(stride[0])),
                  C.c_double(float(stride[1])), tmp, _lib.ptr(tab), _lib.ptr(target), _lib.ptr(weight),
                  _lib.stream_ptr(device))
    return target, weight


def generate_target(joints, joints_vis, heatmap_size, sigma, image_size):
    """uda/dataset/util.py:9-68 - same arguments and numpy return types:
    joints (K,2), joints_vis (K,1), heatmap_size (W,H) -> target (K,H,W) float32, target_weight (K,1)."""
    joints = np.asarray(joints, dtype=np.float64)
    vis = np.asarray(joints_vis, dtype=np.float32)
    t, w = generate_target_batch(joints[None, :, :2], vis.reshape(1, -1, 1), heatmap_size, sigma, image_size)
    return t[0].cpu().numpy(), w[0].cpu().numpy()
