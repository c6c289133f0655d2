"""Load the REAL reference hot-path modules from /root/reference (test infrastructure).

This file is part of the oracle: only ``tests/``, ``oracle/gen_golden.py`` and the
oracle-validation script may import it.  It never ships with the product path and it
only works where the reference tree exists: /root/reference in this container, or the copy of its *.py files
that oracle/ship_reference.py leaves under oracle/_ref/reference/ (git-ignored; travels to the GPU box).

The reference does not import on a modern stack as shipped:
  * ``np.int`` / ``np.float`` were removed from numpy   (regda_4.py:80, regda_7.py:75,3033,3108,3195)
  * ``matplotlib`` / ``webcolors`` are not installed      (regda_4.py:10, regda_7.py:5, keypoint_dataset.py:4)
  * ``uda/model/__init__.py`` pulls in a removed torchvision module (resnet.py:7-8)
so the loader installs shims and loads the four hot-path files *by path*, bypassing the
package ``__init__``.  Nothing is copied: the modules execute from where they lie.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

_CACHE = {}


def reference_root():
    shipped = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "reference")   # oracle/ship_reference.py
    for cand in (os.environ.get("HP_REF_DIR"), "/root/reference", shipped):
        if cand and os.path.isfile(os.path.join(cand, "utils", "keypoint_detection.py")):
            return cand
    return None


def available() -> bool:
    return reference_root() is not None


def _install_shims():
    import numpy as np

    if not hasattr(np, "int"):
        np.int = int  # type: ignore[attr-defined]
    if not hasattr(np, "float"):
        np.float = float  # type: ignore[attr-defined]
    for name in ("matplotlib", "matplotlib.pyplot", "webcolors"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    if "matplotlib" in sys.modules and "matplotlib.pyplot" in sys.modules:
        setattr(sys.modules["matplotlib"], "pyplot", sys.modules["matplotlib.pyplot"])


def _load_by_path(modname: str, path: str):
    spec = importlib.util.spec_from_file_location(modname, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[modname] = mod
    spec.loader.exec_module(mod)
    return mod


def load():
    """Return a namespace with the reference's hot-path symbols."""
    if "ns" in _CACHE:
        return _CACHE["ns"]
    root = reference_root()
    if root is None:
        raise RuntimeError("reference tree not found (set HP_REF_DIR); the oracle restatement "
                           "in oracle/hp_oracle.py and tests/golden/ are the portable checkers")
    sys.dont_write_bytecode = True  # /root/reference is read-only
    _install_shims()
    if root not in sys.path:
        sys.path.insert(0, root)  # for `utils.*` (regda_*.py import utils.gl / utils.net_utils)
    import utils.keypoint_detection as kd  # imports cleanly as shipped

    loss = _load_by_path("_hpref_loss", os.path.join(root, "uda", "model", "loss.py"))
    r4 = _load_by_path("_hpref_regda_4", os.path.join(root, "uda", "model", "regda_4.py"))
    r7 = _load_by_path("_hpref_regda_7", os.path.join(root, "uda", "model", "regda_7.py"))
    du = _load_by_path("_hpref_dataset_util", os.path.join(root, "uda", "dataset", "util.py"))

    ns = types.SimpleNamespace(
        root=root,
        get_max_preds=kd.get_max_preds,
        accuracy=kd.accuracy,
        calc_dists=kd.calc_dists,
        dist_acc=kd.dist_acc,
        find_keypoints_max=kd.find_keypoints_max,
        compute_uv_from_heatmaps=kd.compute_uv_from_heatmaps,
        compute_uv_from_heatmaps2=kd.compute_uv_from_heatmaps2,
        compute_uv_from_heatmaps3=kd.compute_uv_from_heatmaps3,
        JointsMSELoss=loss.JointsMSELoss,
        JointsKLLoss=loss.JointsKLLoss,
        PseudoLabelGenerator=r4.PseudoLabelGenerator,
        RegressionDisparity=r4.RegressionDisparity,
        PseudoLabelGenerator01=r7.PseudoLabelGenerator01,
        PseudoLabelGenerator02=r7.PseudoLabelGenerator02,
        PseudoLabelGenerator03=r7.PseudoLabelGenerator03,
        RegressionDisparityx1=r7.RegressionDisparityx1,
        RegressionDisparityx5=r7.RegressionDisparityx5,
        RegressionDisparityx6=r7.RegressionDisparityx6,
        generate_target=du.generate_target,
        RegressionDisparity2=r4.RegressionDisparity2, RegressionDisparity3=r4.RegressionDisparity3,
        RegressionDisparity4=r4.RegressionDisparity4, RegressionDisparity5=r4.RegressionDisparity5,
        RegressionDisparity6=r4.RegressionDisparity6, RegressionDisparity7=r4.RegressionDisparity7,
        RegressionDisparity8=r4.RegressionDisparity8, RegressionDisparityx2=r7.RegressionDisparityx2,
        RegressionDisparityx3=r7.RegressionDisparityx3, RegressionDisparityx4=r7.RegressionDisparityx4,
        JointsMSELoss0=loss.JointsMSELoss0, JointsKLLoss5=loss.JointsKLLoss5,
    )
    _CACHE["ns"] = ns
    return ns
