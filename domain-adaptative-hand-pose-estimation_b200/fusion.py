"""Multiscale heatmap fusion on the GPU.

The reference does this inline (``train1.py:410-424`` == ``test.py:362-376``) with three
``nn.Upsample(mode='bilinear')`` modules and two elementwise ops; :func:`fuse_multiscale` produces
the same two tensors with one gather+blend kernel each and no intermediate tensors."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib


def _fuse(lo, a_lo, mid, a_mid, hi, a_hi, size):
    lo = _lib.require_cuda(lo.detach(), "fuse(lo)")
    B, K, hl, wl = lo.shape
    H, W = (size, size) if isinstance(size, int) else (int(size[0]), int(size[1]))
    hm = wm = 0
    if mid is not None:
        mid = _lib.require_cuda(mid.detach(), "fuse(mid)")
        if mid.shape[:2] != lo.shape[:2]:
            raise ValueError("fuse: batch/joint dims differ")
        hm, wm = mid.shape[2], mid.shape[3]
    if hi is not None:
        hi = _lib.require_cuda(hi.detach(), "fuse(hi)")
        if tuple(hi.shape) != (B, K, H, W):
            raise ValueError(f"fuse: hi is {tuple(hi.shape)}, expected {(B, K, H, W)}")
    out = torch.empty((B, K, H, W), dtype=torch.float32, device=lo.device)
    with _lib.on_device(lo.device):
        _lib.call("hp_fuse_multiscale", _lib.ptr(lo), hl, wl, C.c_float(a_lo), _lib.ptr(mid), hm, wm, C.c_float(a_mid),
                  _lib.ptr(hi), C.c_float(a_hi), B * K, H, W, _lib.ptr(out), _lib.stream_ptr(lo.device))
    return out


def upsample_bilinear(x, size):
    """``nn.Upsample(size=size, mode='bilinear')(x)`` (align_corners=False)."""
    return _fuse(x, 1.0, None, 0.0, None, 0.0, size)


def fuse_multiscale(y_adv3, y_adv2, size_hi=64, size_mid=32):
    """train1.py:410-424 -> ``(target5, target0)`` with
    ``target5 = 0.5*up_hi(y_adv3) + up_hi(y_adv2)`` and ``target0 = up_mid(y_adv3)`` (inputs detached)."""
    if isinstance(size_hi, int) and isinstance(size_mid, int) and 2 * size_mid == size_hi:
        # both maps from ONE launch (hp_fuse_multiscale_pair; any geometry it does not cover falls back to two inside the call)
        lo = _lib.require_cuda(y_adv3.detach(), "fuse(lo)")
        mid = _lib.require_cuda(y_adv2.detach(), "fuse(mid)")
        if mid.shape[:2] != lo.shape[:2]:
            raise ValueError("fuse: batch/joint dims differ")
        B, K, hl, wl = lo.shape
        target5 = torch.empty((B, K, size_hi, size_hi), dtype=torch.float32, device=lo.device)
        target0 = torch.empty((B, K, size_mid, size_mid), dtype=torch.float32, device=lo.device)
        with _lib.on_device(lo.device):
            _lib.call("hp_fuse_multiscale_pair", _lib.ptr(lo), hl, wl, C.c_float(0.5), _lib.ptr(mid), mid.shape[2], mid.shape[3],
                      C.c_float(1.0), B * K, size_hi, size_hi, _lib.ptr(target5), C.c_float(1.0), size_mid, size_mid,
                      _lib.ptr(target0), _lib.stream_ptr(lo.device))
        return target5, target0
    target5 = _fuse(y_adv3, 0.5, y_adv2, 1.0, None, 0.0, size_hi)
    target0 = _fuse(y_adv3, 1.0, None, 0.0, None, 0.0, size_mid)
    return target5, target0


def fuse_three_scales(lo, mid, hi):
    """BASELINE.json configs[3]: the same rule over three resolutions (e.g. 32/64/128):
    ``0.5*up(lo) + up(mid) + hi`` at the resolution of ``hi``."""
    return _fuse(lo, 0.5, mid, 1.0, hi, 1.0, (hi.shape[2], hi.shape[3]))
