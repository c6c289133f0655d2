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
            accuracy, get_max_preds, h3d.generate_target, stb.generate_target, rhd.generate_target,
            r4.get_max_preds, r7.get_max_preds, r4.RegressionDisparity):
    assert obj.__module__.startswith(ours), (obj, obj.__module__)
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


def test_overlay_rebinds_every_driver_import(tmp_path):
    probe = tmp_path / "driver_probe.py"
    probe.write_text(PROBE)
    p = subprocess.run([sys.executable, "-W", "ignore", os.path.join(ROOT, "hpb200.py"), "--ref", ref_loader.reference_root(),
                        str(probe), "data/H3D", "-t", "Hand3DStudio"], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0 and "overlay-ok" in p.stdout, p.stdout + p.stderr
