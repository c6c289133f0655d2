"""CPU, where the reference tree exists: the overlay rebinding that lets train1.py / test.py run
unchanged - every hot-path name the drivers import resolves to this package."""
import os
import subprocess
import sys

import pytest

from oracle import ref_loader

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present")

PROBE = r'''
# the import block of train1.py:18-32, verbatim names
from uda.model.regda_4 import PseudoLabelGenerator, RegressionDisparity4, RegressionDisparity3
from uda.model.regda_7 import RegressionDisparityx1, RegressionDisparityx5, PseudoLabelGenerator03, \
    PseudoLabelGenerator01, RegressionDisparity, RegressionDisparityx6
from uda.model.loss import JointsKLLoss, JointsMSELoss
from utils.keypoint_detection import accuracy, get_max_preds
import sys, uda.dataset, uda.model.regda_4 as r4, uda.model.regda_7 as r7
h3d, stb, rhd = (sys.modules["uda.dataset." + m] for m in ("hand_3d_studio", "STB", "rendered_hand_pose"))
ours = "domain-adaptative-hand-pose-estimation_b200"
for obj in (PseudoLabelGenerator, RegressionDisparityx1, RegressionDisparityx5, PseudoLabelGenerator03,
            PseudoLabelGenerator01, RegressionDisparity, RegressionDisparityx6, JointsKLLoss, JointsMSELoss,
            accuracy, get_max_preds, r4.get_max_preds, r7.get_max_preds, r4.RegressionDisparity):
    assert obj.__module__.startswith(ours), (obj, obj.__module__)
# the dataset side (DataLoader workers) keeps the reference's numpy generate_target unless --device-targets is given
import os
for m in (h3d, stb, rhd):
    if os.environ.get("HP_EXPECT_DEVICE_TARGETS") == "1":
        assert m.generate_target.__module__.startswith(ours), m.generate_target.__module__
    else:
        assert m.generate_target.__module__ == "uda.dataset.util", m.generate_target.__module__
# nn.Upsample(mode='bilinear') is routed (train1.py:410-417); on CPU tensors it must stay torch's own
import torch, torch.nn as nn
assert nn.Upsample.forward.__module__.startswith(ours)
x = torch.arange(16.0).reshape(1, 1, 4, 4)
assert torch.equal(nn.Upsample(size=8, mode="bilinear")(x), torch.nn.functional.interpolate(x, size=8, mode="bilinear"))
# row f3: the variants the drivers import but never call are rebound too
assert RegressionDisparity3.__module__.startswith(ours) and RegressionDisparity4.__module__.startswith(ours)
from uda.model.loss import JointsMSELoss0, JointsKLLoss5
assert JointsMSELoss0.__module__.startswith(ours) and JointsKLLoss5.__module__.startswith(ours)
# models and everything else stay the reference's own
from uda.model.regda_4 import PoseResNet3
assert PoseResNet3.__module__ == "uda.model.regda_4"
import sys
assert sys.argv[1:] == ["data/H3D", "-t", "Hand3DStudio"], sys.argv
print("overlay-ok")
'''


@pytest.mark.parametrize("device_targets", [False, True])
def test_overlay_rebinds_every_driver_import(tmp_path, device_targets):
    probe = tmp_path / "driver_probe.py"
    probe.write_text(PROBE)
    env = dict(os.environ, HP_EXPECT_DEVICE_TARGETS="1" if device_targets else "0")
    p = subprocess.run([sys.executable, "-W", "ignore", os.path.join(ROOT, "hpb200.py"), "--ref", ref_loader.reference_root()]
                       + (["--device-targets"] if device_targets else []) +
                       [str(probe), "data/H3D", "-t", "Hand3DStudio"], capture_output=True, text=True, timeout=300, env=env)
    assert p.returncode == 0 and "overlay-ok" in p.stdout, p.stdout + p.stderr
