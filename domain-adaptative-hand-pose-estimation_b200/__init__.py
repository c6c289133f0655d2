"""B200-native heatmap keypoint hot path (drop-in for the reference's Python call surface).

Importable as ``importlib.import_module("domain-adaptative-hand-pose-estimation_b200")`` or
through the short alias module ``hpb200`` at the repo root.

Public names mirror the reference:

=====================================  ====================================================
this package                           reference
=====================================  ====================================================
``get_max_preds``, ``accuracy``        utils/keypoint_detection.py:7-35, 63-92
``compute_uv_from_heatmaps{,2,3}``     utils/keypoint_detection.py:139-239 (resize+argmax, soft-argmax)
``JointsMSELoss``, ``JointsKLLoss``    uda/model/loss.py:27-65, 115-158
``PseudoLabelGenerator{,01,02,03}``    uda/model/regda_4.py:17-86, regda_7.py:2956-3201
``RegressionDisparity{,x1,x5,x6}``     uda/model/regda_4.py:89-143, regda_7.py:3206-3632
``generate_target``                    uda/dataset/util.py:9-68
``fuse_multiscale``                    train1.py:410-424 (inline in the reference)
``FusedHeads``                         the same map left unfused: x6 'max' builds it inside the loss kernel
``HeatmapPipeline``                    gen + loss + decode + PCK fused (BASELINE.json metric)
=====================================  ====================================================

Every compute entry point goes through the C-ABI library ``libhp_b200.so`` (``include/hp_b200.h``)
built from ``csrc/``; there is no CPU or eager-PyTorch fallback: a missing library raises.
Submodules are imported lazily so that host-only helpers (``synth``) work without the library.
"""
from __future__ import annotations

import importlib

__version__ = "0.1.0"

_LAZY = {
    "get_max_preds": "keypoint_detection", "accuracy": "keypoint_detection",
    "decode": "keypoint_detection", "pck": "keypoint_detection", "group_accuracy": "keypoint_detection",
    "find_keypoints_max": "keypoint_detection", "compute_uv_from_heatmaps": "keypoint_detection",
    "compute_uv_from_heatmaps2": "keypoint_detection", "compute_uv_from_heatmaps3": "keypoint_detection",
    "JointsMSELoss": "loss", "JointsKLLoss": "loss",
    "PseudoLabelGenerator": "regda", "PseudoLabelGenerator01": "regda",
    "PseudoLabelGenerator02": "regda", "PseudoLabelGenerator03": "regda",
    "RegressionDisparity": "regda", "RegressionDisparityx1": "regda",
    "RegressionDisparityx5": "regda", "RegressionDisparityx6": "regda",
    "RegressionDisparity2": "regda", "RegressionDisparity3": "regda", "RegressionDisparity4": "regda",
    "RegressionDisparity5": "regda", "RegressionDisparity6": "regda", "RegressionDisparity7": "regda",
    "RegressionDisparity8": "regda", "RegressionDisparityx2": "regda", "RegressionDisparityx3": "regda",
    "RegressionDisparityx4": "regda", "JointsMSELoss0": "loss", "JointsKLLoss5": "loss",
    "generate_target": "target", "generate_target_batch": "target", "DeviceTargetCollate": "target",
    "fuse_multiscale": "fusion", "fuse_three_scales": "fusion", "upsample_bilinear": "fusion", "FusedHeads": "fusion",
    "HeatmapPipeline": "pipeline", "PipelineResult": "pipeline",
    "MultiscaleEval": "pipeline",
}


def __getattr__(name):
    if name in _LAZY:
        mod = importlib.import_module("." + _LAZY[name], __name__)
        return getattr(mod, name)
    if name in ("synth", "_lib", "keypoint_detection", "loss", "regda", "target", "fusion",
                "pipeline", "dist", "run", "overlay"):
        return importlib.import_module("." + name, __name__)
    raise AttributeError(name)
