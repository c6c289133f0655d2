"""Reference-shaped facade over :mod:`oracle.hp_oracle` (TEST INFRASTRUCTURE).

Gives the oracle the same names and call signatures as the reference modules (and as the
CUDA package), so one parity case in ``oracle/cases.py`` can be evaluated against the real
reference, the oracle restatement and the B200 path interchangeably."""
from __future__ import annotations

import types

import torch.nn as nn

from . import hp_oracle as O


class JointsMSELoss(nn.Module):                      # uda/model/loss.py:27-65
    def __init__(self, reduction="mean"):
        super().__init__()
        self.reduction = reduction

    def forward(self, output, target, target_weight=None):
        return O.joints_mse_loss(output, target, target_weight, self.reduction)


class JointsKLLoss(nn.Module):                       # uda/model/loss.py:115-158
    def __init__(self, reduction="mean", epsilon=0.0):
        super().__init__()
        self.reduction = reduction
        self.epsilon = epsilon

    def forward(self, output, target, target_weight=None):
        return O.joints_kl_loss(output, target, target_weight, self.reduction, self.epsilon)


class _PLG(nn.Module):
    variant = "base"

    def __init__(self, num_keypoints, height=64, width=64, sigma=2):
        super().__init__()
        self.height, self.width, self.sigma = height, width, sigma

    def forward(self, y):
        return O.pseudo_label(y, self.variant, self.sigma)


class PseudoLabelGenerator(_PLG):                    # uda/model/regda_4.py:17-86
    variant = "base"


class PseudoLabelGenerator02(_PLG):                  # uda/model/regda_7.py:3044-3114 (copy of base)
    variant = "base"


class PseudoLabelGenerator01(_PLG):                  # uda/model/regda_7.py:2956-3039
    variant = "01"

    def __init__(self, num_keypoints, height=16, width=16, sigma=2):
        super().__init__(num_keypoints, height, width, sigma)


class PseudoLabelGenerator03(_PLG):                  # uda/model/regda_7.py:3118-3201
    variant = "03"

    def __init__(self, num_keypoints, height=32, width=32, sigma=2):
        super().__init__(num_keypoints, height, width, sigma)


class _RD(nn.Module):
    variant = "base"

    def __init__(self, pseudo_label_generator, criterion):
        super().__init__()
        self.pseudo_label_generator = pseudo_label_generator
        self.criterion = criterion

    def _run(self, y, y_adv, y_adv2, weight, mode):
        assert mode in ["min", "max"]
        gt, gf = O.ground_maps(self.variant, y.detach(), y_adv2)
        self.ground_truth, self.ground_false = gt, gf
        return self.criterion(y_adv, gt if mode == "min" else gf, weight)


class RegressionDisparity(_RD):                      # uda/model/regda_4.py:89-143
    variant = "base"

    def forward(self, y, y_adv, weight=None, mode="min"):
        return self._run(y, y_adv, None, weight, mode)


class RegressionDisparityx1(_RD):                    # uda/model/regda_7.py:3206-3268
    variant = "x1"

    def forward(self, y, y_adv, weight=None, mode="min"):
        return self._run(y, y_adv, None, weight, mode)


class RegressionDisparityx5(_RD):                    # uda/model/regda_7.py:3485-3561
    variant = "x5"

    def forward(self, y, y_adv, y_adv2, weight=None, mode="min"):
        return self._run(y, y_adv, y_adv2, weight, mode)


class RegressionDisparityx6(_RD):                    # uda/model/regda_7.py:3564-3632
    variant = "x6"

    def forward(self, y, y_adv, y_adv2, weight=None, mode="min"):
        return self._run(y, y_adv, y_adv2, weight, mode)


class RegressionDisparity4(_RD):                     # uda/model/regda_4.py:299-356
    variant = "rd4"

    def forward(self, y, y_adv, weight=None, mode="min"):
        return self._run(y, y_adv, None, weight, mode)


class RegressionDisparityx2(_RD):                    # uda/model/regda_7.py:3272-3337
    variant = "x2"

    def forward(self, y, y_adv, weight=None, mode="min"):
        return self._run(y, y_adv, None, weight, mode)


class RegressionDisparityx3(RegressionDisparityx2):  # uda/model/regda_7.py:3340-3405
    variant = "x3"


class RegressionDisparityx4(_RD):                    # uda/model/regda_7.py:3408-3482
    variant = "x4"

    def forward(self, y, y_adv, weight=None, y_adv2=None, mode="min"):
        return self._run(y, y_adv, None, weight, mode)


class _RDF(nn.Module):
    variant = "rd2"

    def __init__(self, pseudo_label_generator, criterion):
        super().__init__()
        self.pseudo_label_generator = pseudo_label_generator
        self.criterion = criterion

    def _run(self, y, y_adv, label_1, label_2, weight, mode):
        assert mode in ["min", "max"]
        gt, gf = O.ground_maps_fusion(self.variant, y.detach(), label_1.detach(),
                                      None if label_2 is None else label_2.detach())
        self.ground_truth, self.ground_false = gt, gf
        return self.criterion(y_adv, gt if mode == "min" else gf, weight)


def _rdf(variant, two):
    if two:
        def forward(self, y, y_adv, label_1, label_2, weight=None, mode="min"):
            return self._run(y, y_adv, label_1, label_2, weight, mode)
    else:
        def forward(self, y, y_adv, label_1, weight=None, mode="min"):
            return self._run(y, y_adv, label_1, None, weight, mode)
    return type("RegressionDisparity" + variant[2:], (_RDF,), {"variant": variant, "forward": forward})


RegressionDisparity2, RegressionDisparity3 = _rdf("rd2", True), _rdf("rd3", True)      # regda_4.py:145-297
RegressionDisparity5, RegressionDisparity8 = _rdf("rd5", True), _rdf("rd8", True)      # regda_4.py:358-427, 574-645
RegressionDisparity6, RegressionDisparity7 = _rdf("rd6", False), _rdf("rd7", False)    # regda_4.py:429-572


class JointsMSELoss0(nn.Module):                     # uda/model/loss.py:68-112
    def __init__(self, reduction="mean"):
        super().__init__()
        self.reduction = reduction

    def forward(self, output, target, target_weight=None):
        return O.joints_mse_loss0(output, target, target_weight, self.reduction)


class JointsKLLoss5(nn.Module):                      # uda/model/loss.py:160-216
    def __init__(self, reduction="mean", epsilon=0.0):
        super().__init__()
        self.reduction = reduction
        self.epsilon = epsilon

    def forward(self, output, target, target_weight=None):
        return O.joints_kl_loss5(output, target, target_weight, self.reduction, self.epsilon)


def namespace():
    return types.SimpleNamespace(
        get_max_preds=O.get_max_preds, accuracy=O.accuracy, generate_target=O.generate_target,
        find_keypoints_max=O.find_keypoints_max, compute_uv_from_heatmaps=O.compute_uv_from_heatmaps,
        compute_uv_from_heatmaps2=O.compute_uv_from_heatmaps2, compute_uv_from_heatmaps3=O.compute_uv_from_heatmaps3,
        JointsMSELoss=JointsMSELoss, JointsKLLoss=JointsKLLoss,
        PseudoLabelGenerator=PseudoLabelGenerator, PseudoLabelGenerator01=PseudoLabelGenerator01,
        PseudoLabelGenerator02=PseudoLabelGenerator02, PseudoLabelGenerator03=PseudoLabelGenerator03,
        RegressionDisparity=RegressionDisparity, RegressionDisparityx1=RegressionDisparityx1,
        RegressionDisparityx5=RegressionDisparityx5, RegressionDisparityx6=RegressionDisparityx6,
        RegressionDisparity2=RegressionDisparity2, RegressionDisparity3=RegressionDisparity3,
        RegressionDisparity4=RegressionDisparity4, RegressionDisparity5=RegressionDisparity5,
        RegressionDisparity6=RegressionDisparity6, RegressionDisparity7=RegressionDisparity7,
        RegressionDisparity8=RegressionDisparity8, RegressionDisparityx2=RegressionDisparityx2,
        RegressionDisparityx3=RegressionDisparityx3, RegressionDisparityx4=RegressionDisparityx4,
        JointsMSELoss0=JointsMSELoss0, JointsKLLoss5=JointsKLLoss5,
    )
