"""``JointsMSELoss`` / ``JointsKLLoss`` with the reference's constructor and ``forward`` signatures
(``uda/model/loss.py:27-65, 115-158``), each a single fused CUDA pass forward and backward.

Gradients flow to ``output`` only (what ``train1.py:392,433,449`` needs); ``target`` and
``target_weight`` are treated as constants, as they are everywhere in the reference's drivers.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn

from . import _lib


def _flat_weight(w, B, K, device):
    if w is None:
        return None
    w = _lib.require_cuda(w.detach(), "target_weight")
    if w.numel() != B * K:
        raise ValueError(f"target_weight has {w.numel()} elements, expected {B}*{K}")
    return w.reshape(B * K)


class _NoCtx:
    """Stand-in for the autograd context when no gradient is wanted: the forward bodies below run as plain
    functions then (``torch.autograd.Function.apply`` alone costs more host time than the kernel takes)."""
    __slots__ = ("reduction", "epsilon", "cfg", "fz", "w", "heads")

    def save_for_backward(self, *tensors):
        pass


def _wants_grad(t):
    return torch.is_grad_enabled() and isinstance(t, torch.Tensor) and t.requires_grad


class _MSE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, output, target, weight, reduction):
        out = _lib.require_cuda(output.detach(), "JointsMSELoss(output)")
        tgt = _lib.require_cuda(target.detach(), "JointsMSELoss(target)")
        if out.shape != tgt.shape or out.ndim != 4:
            raise ValueError(f"JointsMSELoss: output {tuple(out.shape)} vs target {tuple(tgt.shape)}")
        B, K, H, W = out.shape
        dev = out.device
        w = _flat_weight(weight, B, K, dev)
        per_map = torch.empty((B, K), dtype=torch.float32, device=dev)
        mean = torch.empty((), dtype=torch.float32, device=dev) if reduction == "mean" else None
        with _lib.on_device(dev):
            ws = _lib.workspace(dev, B * K, K)
            _lib.call("hp_mse_fwd", _lib.ptr(out), _lib.ptr(tgt), _lib.ptr(w), B, K, H * W, _lib.ptr(per_map),
                      _lib.ptr(mean), _lib.ptr(ws), _lib.stream_ptr(dev))
        ctx.save_for_backward(out, tgt, w)
        ctx.reduction = reduction
        return mean if reduction == "mean" else per_map

    @staticmethod
    def backward(ctx, grad_out):
        out, tgt, w = ctx.saved_tensors
        B, K, H, W = out.shape
        dev = out.device
        go = grad_out.detach().to(torch.float32).contiguous()
        kind = _lib.GRAD_SCALAR if ctx.reduction == "mean" else _lib.GRAD_PER_MAP
        grad_in = torch.empty_like(out)
        with _lib.on_device(dev):
            _lib.call("hp_mse_bwd", _lib.ptr(out), _lib.ptr(tgt), _lib.ptr(w), _lib.ptr(go), kind, B, K, H * W,
                      _lib.ptr(grad_in), _lib.stream_ptr(dev))
        return grad_in, None, None, None


class _KL(torch.autograd.Function):
    @staticmethod
    def forward(ctx, output, target, weight, reduction, epsilon):
        out = _lib.require_cuda(output.detach(), "JointsKLLoss(output)")
        tgt = _lib.require_cuda(target.detach(), "JointsKLLoss(target)")
        if out.shape != tgt.shape or out.ndim != 4:
            raise ValueError(f"JointsKLLoss: output {tuple(out.shape)} vs target {tuple(tgt.shape)}")
        B, K, H, W = out.shape
        dev = out.device
        w = _flat_weight(weight, B, K, dev)
        # one allocation for the per-map outputs: per_map [B*K] | stats [B*K, 2] (every torch.empty is ~3 us of host time)
        n = B * K
        pack = torch.empty((3 * n,), dtype=torch.float32, device=dev)
        base = pack.data_ptr()
        mean = torch.empty((), dtype=torch.float32, device=dev) if reduction == "mean" else None
        per_sample = torch.empty((B,), dtype=torch.float32, device=dev) if reduction == "none" else None
        with _lib.on_device(dev):
            ws = _lib.workspace(dev, B * K, K)
            _lib.call("hp_kl_fwd", _lib.ptr(out), _lib.ptr(tgt), _lib.ptr(w), C.c_float(epsilon), B, K, H * W,
                      base, _lib.ptr(per_sample), _lib.ptr(mean), base + 4 * n, _lib.ptr(ws),
                      _lib.stream_ptr(dev))
        ctx.save_for_backward(out, tgt, w, pack)
        ctx.reduction = reduction
        ctx.epsilon = float(epsilon)
        return mean if reduction == "mean" else per_sample

    @staticmethod
    def backward(ctx, grad_out):
        out, tgt, w, pack = ctx.saved_tensors
        B, K, H, W = out.shape
        dev = out.device
        go = grad_out.detach()
        if go.dtype != torch.float32 or not go.is_contiguous():
            go = go.to(torch.float32).contiguous()
        kind = _lib.GRAD_SCALAR if ctx.reduction == "mean" else _lib.GRAD_PER_SAMPLE
        grad_in = torch.empty_like(out)
        with _lib.on_device(dev):
            _lib.call("hp_kl_bwd", _lib.ptr(out), _lib.ptr(tgt), _lib.ptr(w), C.c_float(ctx.epsilon), pack.data_ptr() + 4 * B * K,
                      _lib.ptr(go), kind, B, K, H * W, _lib.ptr(grad_in), _lib.stream_ptr(dev))
        return grad_in, None, None, None, None


class JointsMSELoss(nn.Module):
    """uda/model/loss.py:27-65.  ``0.5*(pred-gt)^2*w``; ``'mean'`` over all elements (zero-weight joints
    stay in the denominator), ``'none'`` -> mean over HW -> ``[B,K]``."""

    def __init__(self, reduction="mean"):
        super().__init__()
        self.reduction = reduction

    def forward(self, output, target, target_weight=None):
        if self.reduction not in ("mean", "none"):
            return None                       # the reference falls off the if/elif (loss.py:62-65)
        if not _wants_grad(output):
            return _MSE.forward(_NoCtx(), output, target, target_weight, self.reduction)
        return _MSE.apply(output, target, target_weight, self.reduction)


class JointsKLLoss(nn.Module):
    """uda/model/loss.py:115-158.  KL(q || softmax(pred)) per map, q = (gt+eps)/sum(gt+eps);
    ``'mean'`` over B*K, ``'none'`` -> mean over K -> ``[B]`` (the reference docstring says [B,K];
    its code returns [B], and so does this)."""

    def __init__(self, reduction="mean", epsilon=0.0):
        super().__init__()
        self.reduction = reduction
        self.epsilon = epsilon

    def forward(self, output, target, target_weight=None):
        if self.reduction not in ("mean", "none"):
            return None
        if not _wants_grad(output):
            return _KL.forward(_NoCtx(), output, target, target_weight, self.reduction, float(self.epsilon))
        return _KL.apply(output, target, target_weight, self.reduction, float(self.epsilon))


class _MSE0(torch.autograd.Function):
    @staticmethod
    def forward(ctx, output, target, weight, reduction):
        out = _lib.require_cuda(output.detach(), "JointsMSELoss0(output)")
        tgt = _lib.require_cuda(target.detach(), "JointsMSELoss0(target)")
        if out.shape != tgt.shape or out.ndim != 4:
            raise ValueError(f"JointsMSELoss0: output {tuple(out.shape)} vs target {tuple(tgt.shape)}")
        B, K, H, W = out.shape
        dev = out.device
        w = _flat_weight(weight, B, K, dev)
        per_map = torch.empty((B, K), dtype=torch.float32, device=dev)
        mean = torch.empty((), dtype=torch.float32, device=dev) if reduction == "mean" else None
        with _lib.on_device(dev):
            ws = _lib.workspace(dev, B * K, K)
            _lib.call("hp_mse0_fwd", _lib.ptr(out), _lib.ptr(tgt), _lib.ptr(w), B, K, H * W, _lib.ptr(per_map), _lib.ptr(mean),
                      _lib.ptr(ws), _lib.stream_ptr(dev))
        ctx.save_for_backward(out, tgt, w)
        ctx.reduction = reduction
        return mean if reduction == "mean" else per_map

    @staticmethod
    def backward(ctx, grad_out):
        out, tgt, w = ctx.saved_tensors
        B, K, H, W = out.shape
        dev = out.device
        go = grad_out.detach().to(torch.float32).contiguous()
        kind = _lib.GRAD_SCALAR if ctx.reduction == "mean" else _lib.GRAD_PER_MAP
        grad_in = torch.empty_like(out)
        with _lib.on_device(dev):
            _lib.call("hp_mse0_bwd", _lib.ptr(out), _lib.ptr(tgt), _lib.ptr(w), _lib.ptr(go), kind, B, K, H * W,
                      _lib.ptr(grad_in), _lib.stream_ptr(dev))
        return grad_in, None, None, None


class _KL5(torch.autograd.Function):
    @staticmethod
    def forward(ctx, output, target, reduction, epsilon):
        out = _lib.require_cuda(output.detach(), "JointsKLLoss5(output)")
        tgt = _lib.require_cuda(target.detach(), "JointsKLLoss5(target)")
        if out.shape != tgt.shape or out.ndim != 4:
            raise ValueError(f"JointsKLLoss5: output {tuple(out.shape)} vs target {tuple(tgt.shape)}")
        B, K, H, W = out.shape
        dev = out.device
        scratch = torch.empty((4, B * K), dtype=torch.float32, device=dev)
        per_map = torch.empty((B, K), dtype=torch.float32, device=dev)
        stats = torch.empty((B * K, 2), dtype=torch.float32, device=dev)
        mean = torch.empty((), dtype=torch.float32, device=dev) if reduction == "mean" else None
        per_sample = torch.empty((B,), dtype=torch.float32, device=dev) if reduction == "none" else None
        with _lib.on_device(dev):
            ws = _lib.workspace(dev, B * K, K)
            _lib.call("hp_kl5_fwd", _lib.ptr(out), _lib.ptr(tgt), C.c_float(epsilon), B, K, H * W, _lib.ptr(scratch),
                      _lib.ptr(per_map), _lib.ptr(per_sample), _lib.ptr(mean), _lib.ptr(stats), _lib.ptr(ws),
                      _lib.stream_ptr(dev))
        ctx.save_for_backward(out, tgt, scratch, stats)
        ctx.reduction = reduction
        ctx.epsilon = float(epsilon)
        return mean if reduction == "mean" else per_sample

    @staticmethod
    def backward(ctx, grad_out):
        out, tgt, scratch, stats = ctx.saved_tensors
        B, K, H, W = out.shape
        dev = out.device
        go = grad_out.detach().to(torch.float32).contiguous()
        kind = _lib.GRAD_SCALAR if ctx.reduction == "mean" else _lib.GRAD_PER_SAMPLE
        grad_in = torch.empty_like(out)
        with _lib.on_device(dev):
            _lib.call("hp_kl5_bwd", _lib.ptr(out), _lib.ptr(tgt), C.c_float(ctx.epsilon), _lib.ptr(scratch[3]),
                      _lib.ptr(stats), _lib.ptr(go), kind, B, K, H * W, _lib.ptr(grad_in), _lib.stream_ptr(dev))
        return grad_in, None, None, None


class JointsMSELoss0(nn.Module):
    """uda/model/loss.py:68-112: prediction and label are each shifted by 1e-7 and normalised to sum 1 per map, then
    ``0.5*(p-t)^2*w``.  One block-per-map kernel each way (``hp_mse0_fwd`` / ``hp_mse0_bwd``, csrc/hp_loss_variants.cu):
    the normalised maps exist in registers only and autograd sees no intermediate tensor."""

    def __init__(self, reduction="mean"):
        super().__init__()
        self.reduction = reduction

    def forward(self, output, target, target_weight=None):
        if self.reduction not in ("mean", "none"):
            return None                       # the reference falls off the if/elif (loss.py:109-112)
        if not _wants_grad(output):
            return _MSE0.forward(_NoCtx(), output, target, target_weight, self.reduction)
        return _MSE0.apply(output, target, target_weight, self.reduction)


class JointsKLLoss5(nn.Module):
    """uda/model/loss.py:160-216: both tensors are rescaled per map by ``w5`` (a detached overlap score of prediction and
    label, normalised by its global maximum), then the KL loss WITHOUT target weights (the reference ignores them).
    ``hp_kl5_fwd`` = per-map statistics, one-block scale, per-map KL; ``hp_kl5_bwd`` = the gradient through the scaled
    logits (csrc/hp_loss_variants.cu)."""

    def __init__(self, reduction="mean", epsilon=0.):
        super().__init__()
        self.reduction = reduction
        self.epsilon = epsilon

    def forward(self, output, target, target_weight=None):
        if self.reduction not in ("mean", "none"):
            return None                       # loss.py:213-216
        if not _wants_grad(output):
            return _KL5.forward(_NoCtx(), output, target, self.reduction, float(self.epsilon))
        return _KL5.apply(output, target, self.reduction, float(self.epsilon))
