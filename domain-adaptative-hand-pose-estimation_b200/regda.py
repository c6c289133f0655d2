"""Pseudo-label generators and regression-disparity modules with the reference's signatures.

=============================  ==========================================  ===================
class                          reference                                   pseudo-label grid
=============================  ==========================================  ===================
``PseudoLabelGenerator``       uda/model/regda_4.py:17-86                  H x W of ``y``
``PseudoLabelGenerator02``     uda/model/regda_7.py:3044-3114 (copy)       H x W of ``y``
``PseudoLabelGenerator03``     uda/model/regda_7.py:3118-3201              32 x 32, centre = pred/2, 9x9 patch
``PseudoLabelGenerator01``     uda/model/regda_7.py:2956-3039              16 x 16, centre = pred/4, 7x7 patch
``RegressionDisparity``        uda/model/regda_4.py:89-143                 gf = clip(sum_{j!=k} gt_j)
``RegressionDisparityx1``      uda/model/regda_7.py:3206-3268              gf = clip(1 - 10 gt)
``RegressionDisparityx5``      uda/model/regda_7.py:3485-3561              + fused map, per-map max normalise
``RegressionDisparityx6``      uda/model/regda_7.py:3564-3632              clip(sum gt) based, + fused map, normalise
=============================  ==========================================  ===================

Unlike the reference nothing leaves the GPU: ``y`` is decoded by a CUDA kernel, and when the
criterion is this package's ``JointsKLLoss`` (what ``train1.py:135-137`` wires) the whole disparity
term - pseudo-label, ground-false recipe, normalisation, KL - is one fused kernel that never writes
gt / gf.  ``.ground_truth`` / ``.ground_false`` stay available (materialised on first access).
Any other criterion gets materialised maps and is called as in the reference.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from .fusion import FusedHeads, LazyUpsample
from .loss import JointsKLLoss, _NoCtx, _wants_grad

_VARIANT_CODE = {"base": _lib.RD_BASE, "x1": _lib.RD_X1, "x5": _lib.RD_X5, "x6": _lib.RD_X6, "rd4": _lib.RD_RD4}


class _PLGBase(nn.Module):
    #: (output side or None, shift applied to the decoded coordinate, tmp_size / sigma, C-ABI kind)
    _spec = (None, 0, 3, _lib.PLG_BASE)

    def __init__(self, num_keypoints, height=64, width=64, sigma=2):
        super().__init__()
        self.num_keypoints = num_keypoints
        self.height = height
        self.width = width
        self.sigma = sigma
        # regda_4.py:74 - kept for API parity; the kernels use sum-minus-self instead of the sgemm
        self.false_matrix = 1.0 - np.eye(num_keypoints, dtype=np.float32)

    # -- geometry --------------------------------------------------------------------------
    def grid(self, H, W):
        side, shift, tmp_factor, kind = self._spec
        oh, ow = (H, W) if side is None else (side, side)
        return oh, ow, shift, _lib.integer_tmp(self.sigma * tmp_factor), kind

    def _check_input(self, y):
        y = _lib.require_cuda(y.detach(), type(self).__name__ + "(y)")
        if y.ndim != 4:
            raise ValueError("y must be [B,K,H,W]")
        B, K, H, W = y.shape
        if K > _lib.MAX_K:
            raise ValueError(f"K={K} exceeds HP_MAX_K={_lib.MAX_K}")
        oh, ow, shift, tmp, kind = self.grid(H, W)
        if ((H - 1) >> shift) >= oh or ((W - 1) >> shift) >= ow:
            # the reference indexes its look-up table out of bounds here (IndexError)
            raise IndexError(f"decoded coordinates of a {H}x{W} map do not fit the {oh}x{ow} pseudo-label grid")
        return y, (B, K, H, W), (oh, ow, shift, tmp, kind)

    def forward(self, y):
        """-> (ground_truth, ground_false) float32 [B,K,oh,ow] on ``y.device``."""
        y, (B, K, H, W), (oh, ow, shift, tmp, kind) = self._check_input(y)
        dev = y.device
        gt = torch.empty((B, K, oh, ow), dtype=torch.float32, device=dev)
        gf = torch.empty_like(gt)
        centres = torch.empty((B * K, 2), dtype=torch.int32, device=dev)
        with _lib.on_device(dev):
            tab = _lib.gaussian_table(self.sigma, tmp, dev)
            _lib.call("hp_pseudo_label", _lib.ptr(y), B, K, H, W, kind, oh, ow, shift, tmp, _lib.ptr(tab), _lib.ptr(gt),
                      _lib.ptr(gf), _lib.ptr(centres), _lib.stream_ptr(dev))
        return gt, gf


class PseudoLabelGenerator(_PLGBase):
    _spec = (None, 0, 3, _lib.PLG_BASE)


class PseudoLabelGenerator02(_PLGBase):
    _spec = (None, 0, 3, _lib.PLG_BASE)


class PseudoLabelGenerator03(_PLGBase):
    _spec = (32, 1, 2, _lib.PLG_ONE_MINUS)

    def __init__(self, num_keypoints, height=32, width=32, sigma=2):
        super().__init__(num_keypoints, height, width, sigma)


class PseudoLabelGenerator01(_PLGBase):
    _spec = (16, 2, 1.5, _lib.PLG_ONE_MINUS)

    def __init__(self, num_keypoints, height=16, width=16, sigma=2):
        super().__init__(num_keypoints, height, width, sigma)


# ------------------------------------------------------------------------------------------

class _RegDisp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y_adv, y, fused, weight, variant, mode, epsilon, reduction, plg, holder):
        yd, (B, K, H, W), (oh, ow, shift, tmp, _) = plg._check_input(y)
        adv = _lib.require_cuda(y_adv.detach(), "RegressionDisparity(y_adv)")
        if tuple(adv.shape) != (B, K, oh, ow):
            raise ValueError(f"y_adv is {tuple(adv.shape)}, the pseudo-label grid is {(B, K, oh, ow)}")
        dev = adv.device
        fz = heads = None
        if isinstance(fused, LazyUpsample):      # (the overlay's nn.Upsample route; x5 'max' reads target0 as a map)
            fused = None if mode == _lib.MODE_MIN else fused.materialise()
        if isinstance(fused, FusedHeads):
            # y_adv2 given as its two heads (train1.py:410-424 unfused): x6 'max' on the 64x64 grid builds the map inside the loss
            # kernel; 'min' never reads y_adv2 (regda_7.py:3614-3631); everything else gets the materialised map
            if tuple(fused.shape) != (B, K, oh, ow):
                raise ValueError(f"y_adv2 is {tuple(fused.shape)}, expected {(B, K, oh, ow)}")
            if mode == _lib.MODE_MIN:
                fused = None
            elif (variant == _lib.RD_X6 and fused.in_kernel() and (H, W) == (oh, ow) and tmp <= 6 and K <= 32
                  and fused.lo.device == dev and fused.mid.device == dev and adv.data_ptr() % 16 == 0):
                heads, fused = fused, None
            else:
                fused = fused.materialise()
        if fused is not None:
            fz = _lib.require_cuda(fused.detach(), "RegressionDisparity(y_adv2)")
            if tuple(fz.shape) != (B, K, oh, ow):
                raise ValueError(f"y_adv2 is {tuple(fz.shape)}, expected {(B, K, oh, ow)}")
        w = None
        if weight is not None:
            w = _lib.require_cuda(weight.detach(), "weight").reshape(-1)
            if w.numel() != B * K:
                raise ValueError(f"weight has {w.numel()} elements, expected {B}*{K}")
        # one allocation for the per-map outputs: per_map [B*K] | stats [B*K, 3] | centres [B*K, 2] (int32 bits) - 4-byte words
        n = B * K
        pack = torch.empty((6 * n,), dtype=torch.float32, device=dev)
        base = pack.data_ptr()
        p_per_map, p_stats, p_centres = base, base + 4 * n, base + 16 * n
        mean = torch.empty((), dtype=torch.float32, device=dev) if reduction == "mean" else None
        per_sample = torch.empty((B,), dtype=torch.float32, device=dev) if reduction == "none" else None
        with _lib.on_device(dev):
            tab = _lib.gaussian_table(plg.sigma, tmp, dev)
            ws = _lib.workspace(dev, B * K, K)
            if heads is not None:
                _lib.call("hp_regdisp_fwd_heads", _lib.ptr(yd), _lib.ptr(adv), _lib.ptr(heads.lo), 16, 16, C.c_float(heads.a_lo),
                          _lib.ptr(heads.mid), 32, 32, C.c_float(heads.a_mid), _lib.ptr(w), variant, mode, C.c_float(epsilon),
                          B, K, H, W, oh, ow, tmp, _lib.ptr(tab), p_per_map, _lib.ptr(per_sample), _lib.ptr(mean), p_stats,
                          p_centres, _lib.ptr(ws), _lib.stream_ptr(dev))
            else:
                _lib.call("hp_regdisp_fwd", _lib.ptr(yd), _lib.ptr(adv), _lib.ptr(fz), _lib.ptr(w), variant, mode,
                          C.c_float(epsilon), B, K, H, W, oh, ow, shift, tmp, _lib.ptr(tab), p_per_map,
                          _lib.ptr(per_sample), _lib.ptr(mean), p_stats, p_centres, _lib.ptr(ws),
                          _lib.stream_ptr(dev))
        ctx.save_for_backward(adv, pack, tab)
        ctx.fz, ctx.w, ctx.heads = fz, w, heads
        ctx.cfg = (variant, mode, float(epsilon), reduction, B, K, oh, ow, tmp)
        holder._remember(variant, heads if heads is not None else fz, pack, tab, (B, K, oh, ow, tmp))
        return mean if reduction == "mean" else per_sample

    @staticmethod
    def backward(ctx, grad_out):
        adv, pack, tab = ctx.saved_tensors
        variant, mode, eps, reduction, B, K, oh, ow, tmp = ctx.cfg
        dev = adv.device
        base = pack.data_ptr()
        p_stats, p_centres = base + 4 * B * K, base + 16 * B * K
        go = grad_out.detach()
        if go.dtype != torch.float32 or not go.is_contiguous():
            go = go.to(torch.float32).contiguous()
        kind = _lib.GRAD_SCALAR if reduction == "mean" else _lib.GRAD_PER_SAMPLE
        grad_in = torch.empty_like(adv)
        with _lib.on_device(dev):
            heads = getattr(ctx, "heads", None)
            if heads is not None:
                _lib.call("hp_regdisp_bwd_heads", _lib.ptr(adv), _lib.ptr(heads.lo), 16, 16, C.c_float(heads.a_lo),
                          _lib.ptr(heads.mid), 32, 32, C.c_float(heads.a_mid), _lib.ptr(ctx.w), variant, mode, C.c_float(eps),
                          B, K, oh, ow, tmp, _lib.ptr(tab), p_centres, p_stats, _lib.ptr(go), kind, _lib.ptr(grad_in),
                          _lib.stream_ptr(dev))
            else:
                _lib.call("hp_regdisp_bwd", _lib.ptr(adv), _lib.ptr(ctx.fz), _lib.ptr(ctx.w), variant, mode, C.c_float(eps),
                          B, K, oh, ow, tmp, _lib.ptr(tab), p_centres, p_stats, _lib.ptr(go), kind,
                          _lib.ptr(grad_in), _lib.stream_ptr(dev))
        return (grad_in,) + (None,) * 9


class _RDBase(nn.Module):
    _variant = "base"

    def __init__(self, pseudo_label_generator, criterion: nn.Module):
        super().__init__()
        self.criterion = criterion
        self.pseudo_label_generator = pseudo_label_generator
        self._lazy = None
        self._gt = None
        self._gf = None

    # -- lazily materialised attributes the reference sets eagerly (regda_4.py:136-137) ------
    def _remember(self, variant, fused, centres, tab, dims):
        # (object.__setattr__: nn.Module.__setattr__ costs ~3 us per assignment - it was a third of a forward call's host time)
        object.__setattr__(self, "_lazy", (variant, fused, centres, tab, dims))
        object.__setattr__(self, "_gt", None)
        object.__setattr__(self, "_gf", None)

    def _materialise(self):
        if self._gt is None:
            if self._lazy is None:
                raise AttributeError("ground_truth / ground_false exist only after a forward call")
            variant, fused, centres, tab, (B, K, oh, ow, tmp) = self._lazy
            if isinstance(fused, FusedHeads):
                fused = fused.materialise()
            dev = centres.device
            gt = torch.empty((B, K, oh, ow), dtype=torch.float32, device=dev)
            gf = torch.empty_like(gt)
            with _lib.on_device(dev):
                # `centres` is either an int32 [B*K, 2] tensor or the forward's packed float32 output buffer
                # (per_map | stats | centres), whose centres start 16 * B * K bytes in
                c_ptr = centres.data_ptr() + (16 * B * K if centres.dtype == torch.float32 else 0)
                _lib.call("hp_regdisp_materialize", _lib.ptr(fused), variant, B, K, oh, ow, tmp, _lib.ptr(tab),
                          c_ptr, _lib.ptr(gt), _lib.ptr(gf), _lib.stream_ptr(dev))
            object.__setattr__(self, "_gt", gt)
            object.__setattr__(self, "_gf", gf)
        return self._gt, self._gf

    @property
    def ground_truth(self):
        return self._materialise()[0]

    @property
    def ground_false(self):
        return self._materialise()[1]

    # -----------------------------------------------------------------------------------------
    def _run(self, y, y_adv, y_adv2, weight, mode):
        assert mode in ["min", "max"]
        variant = _VARIANT_CODE[self._variant]
        plg = self.pseudo_label_generator
        if not isinstance(plg, _PLGBase):
            raise TypeError("pseudo_label_generator must be one of this package's PseudoLabelGenerator classes")
        mode_code = _lib.MODE_MIN if mode == "min" else _lib.MODE_MAX
        crit = self.criterion
        if isinstance(crit, JointsKLLoss) and crit.reduction in ("mean", "none"):
            args = (y_adv, y, y_adv2, weight, variant, mode_code, float(crit.epsilon), crit.reduction, plg, self)
            if not _wants_grad(y_adv):
                return _RegDisp.forward(_NoCtx(), *args)
            return _RegDisp.apply(*args)
        # foreign criterion: materialise the maps on the GPU and call it like the reference does
        from .keypoint_detection import decode
        yd, (B, K, H, W), (oh, ow, shift, tmp, _) = plg._check_input(y)
        dev = yd.device
        if isinstance(y_adv2, (FusedHeads, LazyUpsample)):
            y_adv2 = y_adv2.materialise()
        fz = None if y_adv2 is None else _lib.require_cuda(y_adv2.detach(), "y_adv2")
        preds, _ = decode(yd)
        centres = (preds.reshape(-1, 2).to(torch.int32) >> shift).contiguous()
        with _lib.on_device(dev):
            tab = _lib.gaussian_table(plg.sigma, tmp, dev)
        self._remember(variant, fz, centres, tab, (B, K, oh, ow, tmp))
        gt, gf = self._materialise()
        return crit(y_adv, gt if mode == "min" else gf, weight)


class RegressionDisparity(_RDBase):
    """uda/model/regda_4.py:89-143: ``forward(y, y_adv, weight=None, mode='min')``."""
    _variant = "base"

    def forward(self, y, y_adv, weight=None, mode="min"):
        return self._run(y, y_adv, None, weight, mode)


class RegressionDisparityx1(_RDBase):
    """uda/model/regda_7.py:3206-3268 (16x16 head): ``forward(y, y_adv, weight=None, mode='min')``."""
    _variant = "x1"

    def forward(self, y, y_adv, weight=None, mode="min"):
        return self._run(y, y_adv, None, weight, mode)


class RegressionDisparityx5(_RDBase):
    """uda/model/regda_7.py:3485-3561 (32x32 head): ``forward(y, y_adv, y_adv2, weight=None, mode='min')``."""
    _variant = "x5"

    def forward(self, y, y_adv, y_adv2, weight=None, mode="min"):
        return self._run(y, y_adv, y_adv2, weight, mode)


class RegressionDisparityx6(_RDBase):
    """uda/model/regda_7.py:3564-3632 (64x64 head): ``forward(y, y_adv, y_adv2, weight=None, mode='min')``."""
    _variant = "x6"

    def forward(self, y, y_adv, y_adv2, weight=None, mode="min"):
        return self._run(y, y_adv, y_adv2, weight, mode)


# ------------------------------------------------------------------------------------------
# the remaining disparity variants (dead code in the reference's drivers, imported at
# train1.py:19).  Three of them are existing kernel recipes under other names; the label-fusing ones compose the
# CUDA pseudo-label and loss kernels with a few small elementwise torch ops on the GPU (nothing leaves the device).
# ------------------------------------------------------------------------------------------

class RegressionDisparity4(_RDBase):
    """uda/model/regda_4.py:299-356: ``gf = clip(clip(sum_k gt) - 10 gt)`` - the x6 recipe without a fused map and
    without the per-map normalisation.  One fused kernel (variant HP_RD_RD4)."""
    _variant = "rd4"

    def forward(self, y, y_adv, weight=None, mode="min"):
        return self._run(y, y_adv, None, weight, mode)


class RegressionDisparityx2(_RDBase):
    """uda/model/regda_7.py:3272-3337: ``gf = clip(1 - 10 gt)`` (its ``label_p`` is computed and never used)."""
    _variant = "x1"

    def forward(self, y, y_adv, weight=None, mode="min"):
        return self._run(y, y_adv, None, weight, mode)


class RegressionDisparityx3(RegressionDisparityx2):
    """uda/model/regda_7.py:3340-3405: the same recipe as x2."""


class RegressionDisparityx4(_RDBase):
    """uda/model/regda_7.py:3408-3482: ``gf = PLG.gf / max(PLG.gf)`` per map; with the generators that produce
    ``clip(1 - 10 gt)`` the maximum is 1, i.e. the x1 recipe.  A tensor ``y_adv2`` makes the reference raise
    (``if y_adv2:`` on a multi-element tensor), and so does this."""
    _variant = "x1"

    def forward(self, y, y_adv, weight=None, y_adv2=None, mode="min"):
        if y_adv2 is not None and bool(y_adv2):   # raises for multi-element tensors, like the reference
            raise NameError("RegressionDisparityx4: ground_false is undefined when y_adv2 is given (regda_7.py:3463-3466)")
        if self.pseudo_label_generator._spec[3] != _lib.PLG_ONE_MINUS:
            raise NotImplementedError("RegressionDisparityx4 is built for PseudoLabelGenerator01 / 03")
        return self._run(y, y_adv, None, weight, mode)


class _RDLabelFusion(nn.Module):
    """RegressionDisparity2/3/5/6/7/8 (uda/model/regda_4.py:145-645): ``gf = clip(label_p - 10 gt)`` where the per-sample
    map ``label_p`` fuses the summed pseudo-labels of ``y`` and of one or two extra predictions.

    Three decode launches (one per prediction) + ONE kernel (``hp_label_fusion``, csrc/hp_variants.cu: a block per sample
    rebuilds the per-sample sums from the 3K centres, applies the variant's rule and writes gt / gf), then the criterion on
    the materialised target - no elementwise ATen chain, no [B,K,H,W] temporaries besides the two maps the reference itself
    exposes as ``ground_truth`` / ``ground_false``."""

    _rule = 0

    def __init__(self, pseudo_label_generator, criterion: nn.Module):
        super().__init__()
        self.criterion = criterion
        self.pseudo_label_generator = pseudo_label_generator
        self.reset()

    def reset(self):
        self.label_x = 0.

    def updata(self, label_x):          # (sic) regda_4.py:197
        self.label_x = label_x

    def _run(self, y, y_adv, label_1, label_2, weight, mode):
        assert mode in ["min", "max"]
        from .keypoint_detection import decode
        plg = self.pseudo_label_generator
        if not isinstance(plg, _PLGBase):
            raise TypeError("pseudo_label_generator must be one of this package's PseudoLabelGenerator classes")
        yd, (B, K, H, W), (oh, ow, shift, tmp, _) = plg._check_input(y)
        dev = yd.device

        def centres(t):
            td = plg._check_input(t)[0]
            if tuple(td.shape) != (B, K, H, W):
                raise ValueError(f"label is {tuple(td.shape)}, expected {(B, K, H, W)}")
            preds, _ = decode(td)                                   # masked (x, y) like get_max_preds (regda_4.py:79-81)
            return (preds.reshape(-1, 2).to(torch.int32) >> shift).contiguous()

        c0, c1 = centres(y), centres(label_1)
        c2 = centres(label_2) if label_2 is not None else None
        gt = torch.empty((B, K, oh, ow), dtype=torch.float32, device=dev)
        gf = torch.empty_like(gt)
        with _lib.on_device(dev):
            tab = _lib.gaussian_table(plg.sigma, tmp, dev)
            _lib.call("hp_label_fusion", _lib.ptr(c0), _lib.ptr(c1), _lib.ptr(c2), self._rule, B, K, oh, ow, tmp,
                      _lib.ptr(tab), _lib.ptr(gt), _lib.ptr(gf), None, _lib.stream_ptr(dev))
        self.ground_truth, self.ground_false = gt, gf
        return self.criterion(y_adv, gt if mode == "min" else gf, weight)


class _RDLabelFusion2(_RDLabelFusion):
    def forward(self, y, y_adv, label_1, label_2, weight=None, mode="min"):
        return self._run(y, y_adv, label_1, label_2, weight, mode)


class _RDLabelFusion1(_RDLabelFusion):
    def forward(self, y, y_adv, label_1, weight=None, mode="min"):
        return self._run(y, y_adv, label_1, None, weight, mode)


class RegressionDisparity2(_RDLabelFusion2):     # regda_4.py:145-220: lp = (sum gt1 + sum gt + sum gt2) / max
    _rule = 2


class RegressionDisparity3(_RDLabelFusion2):     # regda_4.py:222-297: lp = (c(sum gt) + c(sum gt1) + c(sum gt2)) / max
    _rule = 3


class RegressionDisparity5(_RDLabelFusion2):     # regda_4.py:358-427: lp = c(p1 + c(p2 - p1) + c(p3 - p1))
    _rule = 5


class RegressionDisparity6(_RDLabelFusion1):     # regda_4.py:429-495: lp = c(p1 + c(p2 - p1))
    _rule = 6


class RegressionDisparity7(_RDLabelFusion1):     # regda_4.py:497-572: lp = (c(sum gt1) + c(sum gt)) / max
    _rule = 7


class RegressionDisparity8(_RDLabelFusion2):     # regda_4.py:574-645: lp = c(p1 + c(sum c(gt1 - gt)) + c(sum c(gt2 - gt)))
    _rule = 8
