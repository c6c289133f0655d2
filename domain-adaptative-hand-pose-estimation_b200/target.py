"""Gaussian target generation on the GPU (``uda/dataset/util.py:9-68``).

``generate_target`` keeps the reference's per-sample numpy signature (the datasets call it from
``__getitem__``: ``hand_3d_studio.py:102``, ``STB.py:150``, ``rendered_hand_pose.py:83``);
``generate_target_batch`` is the B200-native form: keypoints of a whole batch in, targets resident
on the device out (no per-sample numpy, no host->device copy of labels)."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.utils.data

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
    if torch.utils.data.get_worker_info() is not None:
        # a forked DataLoader worker cannot use the parent's CUDA context (and a spawned one would pay a
        # context + one launch and one synchronous copy per sample): say so instead of crashing inside CUDA
        raise RuntimeError("hpb200.generate_target was called inside a DataLoader worker process; run the loader "
                           "with num_workers=0, keep the reference's numpy generate_target on the dataset side (the "
                           "overlay's default), or generate targets per batch with generate_target_batch / "
                           "DeviceTargetCollate")
    joints = np.asarray(joints, dtype=np.float64)
    vis = np.asarray(joints_vis, dtype=np.float32)
    t, w = generate_target_batch(joints[None, :, :2], vis.reshape(1, -1, 1), heatmap_size, sigma, image_size)
    return t[0].cpu().numpy(), w[0].cpu().numpy()


class DeviceTargetCollate:
    """Loader-side use of the batched generator (``hand_3d_studio.py:98-104``): a ``collate_fn``
    for datasets that return ``(image, target, target_weight, meta)`` with ``meta['keypoint2d']`` in image pixels.

    It collates as usual, then - in the MAIN process, on the batch - replaces ``target`` / ``target_weight`` by
    :func:`generate_target_batch` evaluated on the collated keypoints (one kernel for the batch instead of B numpy
    calls and a host->device copy of B*K maps).  ``visible`` comes from ``meta['visible']`` when the dataset provides
    it, else all ones (what ``hand_3d_studio.py:98-99`` uses).  Because a collate_fn runs inside the worker when
    ``num_workers > 0``, wrap the LOADER instead in that case: ``DeviceTargetCollate.wrap(loader)`` yields batches
    with the targets regenerated after they left the workers."""

    def __init__(self, heatmap_size, sigma, image_size, device=None, base_collate=None):
        self.heatmap_size, self.sigma, self.image_size = tuple(heatmap_size), sigma, tuple(image_size)
        self.device = device
        self.base = base_collate or torch.utils.data.default_collate

    def regenerate(self, batch):
        image, _, _, meta = batch
        kp = torch.as_tensor(meta["keypoint2d"])[..., :2]
        vis = meta.get("visible")
        if vis is None:
            vis = torch.ones(kp.shape[:2], dtype=torch.float32)
        target, weight = generate_target_batch(kp, vis, self.heatmap_size, self.sigma, self.image_size, self.device)
        return image, target, weight, meta

    def __call__(self, samples):
        batch = self.base(samples)
        if torch.utils.data.get_worker_info() is not None:
            return batch                       # inside a worker: leave the labels alone, see wrap()
        return self.regenerate(batch)

    def wrap(self, loader):
        for batch in loader:
            yield self.regenerate(batch)
