"""Launcher plugin for the overlay (``python hpb200.py --ref DIR --plugin tests/train1_synthetic_plugin.py train1.py ...``).

SURVEY.md row f1 asks for real ``train()`` iterations THROUGH THE UNCHANGED ``train1.py``; datasets and ImageNet weights are
not reachable offline, so this file - run by the launcher after the overlay is installed and before the driver - registers
two names the driver's own flags can select (``-s SyntheticHands -t SyntheticHands -a tinynet``):

* ``uda.dataset.SyntheticHands``: a ``Hand21KeypointDataset`` (the reference's own base class: ``group_accuracy``,
  ``keypoints_group``, ``num_keypoints`` come from it) whose samples are seeded noise images with seeded keypoints; labels
  are produced by whatever ``uda.dataset.util.generate_target`` is bound to (the reference's numpy version by default, the
  CUDA one under ``--device-targets``) - exactly what ``hand_3d_studio.py:98-104`` does;
* ``uda.model.tinynet``: a 5-conv stride-32 stand-in for ``resnet101`` with the ``out_features`` attribute
  ``Upsampling`` reads (``train1.py:123-125``).

It also counts the C-ABI entry points the run goes through and writes them to ``$HP_OVERLAY_STATS`` at exit, so the test
can assert that the CUDA kernels - not a fallback - did the work."""
import atexit
import importlib
import json
import os

import numpy as np
import torch
import torch.nn as nn

import uda.dataset as datasets
import uda.dataset.util as dataset_util
import uda.model as models
from uda.dataset.keypoint_dataset import Hand21KeypointDataset


class SyntheticHands(Hand21KeypointDataset):
    def __init__(self, root, split="train", task="all", download=False, **kwargs):
        kwargs.pop("transforms", None)
        super().__init__(root, list(range(16 if split == "train" else 8)), **kwargs)
        self.split = split

    def __getitem__(self, index):
        rs = np.random.RandomState(7000 + index + (0 if self.split == "train" else 500))
        w, h = self.image_size
        image = torch.from_numpy(rs.standard_normal((3, h, w)).astype(np.float32))
        keypoint2d = rs.uniform(16.0, w - 16.0, size=(self.num_keypoints, 2))
        visible = np.ones((self.num_keypoints, 1), dtype=np.float32)
        target, target_weight = dataset_util.generate_target(keypoint2d, visible, self.heatmap_size, self.sigma,
                                                             self.image_size)
        meta = {"image": f"synthetic_{self.split}_{index}", "keypoint2d": keypoint2d,
                "keypoint3d": np.zeros((self.num_keypoints, 3)), "image_ema": image}
        return image, torch.from_numpy(np.asarray(target)), torch.from_numpy(np.asarray(target_weight)), meta


class TinyBackbone(nn.Module):
    """256 x 256 -> 8 x 8 x 64 (stride 32 like ResNet), so the reference's Upsampling lands on 64 x 64."""

    def __init__(self):
        super().__init__()
        chans = (3, 16, 32, 32, 64, 64)
        layers = []
        for cin, cout in zip(chans[:-1], chans[1:]):
            layers += [nn.Conv2d(cin, cout, 3, 2, 1), nn.BatchNorm2d(cout), nn.ReLU(inplace=True)]
        self.body = nn.Sequential(*layers)
        self.out_features = chans[-1]

    def forward(self, x):
        return self.body(x)


def tinynet(pretrained=False, **kwargs):
    return TinyBackbone()


datasets.SyntheticHands = SyntheticHands
models.tinynet = tinynet

# ---- which C entry points did the run go through? ---------------------------------------------------------------
_stats_path = os.environ.get("HP_OVERLAY_STATS")
if _stats_path:
    _lib = importlib.import_module("domain-adaptative-hand-pose-estimation_b200._lib")
    _counts = {}
    _call = _lib.call

    def _counting_call(name, *args):
        _counts[name] = _counts.get(name, 0) + 1
        return _call(name, *args)

    _lib.call = _counting_call
    for _m in ("keypoint_detection", "loss", "regda", "fusion", "target", "pipeline"):
        _mod = importlib.import_module("domain-adaptative-hand-pose-estimation_b200." + _m)
        if getattr(_mod, "_lib", None) is _lib:
            pass                                  # modules call _lib.call through the module attribute: already counted

    def _dump():
        with open(_stats_path, "w") as f:
            json.dump(_counts, f)
    atexit.register(_dump)
