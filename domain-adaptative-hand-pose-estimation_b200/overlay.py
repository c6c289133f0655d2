"""Run the reference's own drivers (``train1.py`` / ``test.py``) UNCHANGED on top of the B200 path.

    python hpb200.py --ref /path/to/reference train1.py data/H3D -t Hand3DStudio ...
    python hpb200.py --ref /path/to/reference test.py  data/H3D -t Hand3DStudio --checkpoint ...
    options before the driver: --verbose | --eager-upsample (see install_upsample_route) | --device-targets (rebind the per-sample generate_target too; needs
    --workers 0) | --plugin FILE.py (run FILE after the overlay is installed and before the driver, e.g. to register a
    synthetic dataset class in ``uda.dataset`` or a stand-in backbone in ``uda.model`` under a name the driver's
    ``-s/-t/-a`` flags can select)

The reference has no plugin registry: its hot-path callables are plain names imported at
``train1.py:18-32`` from modules that also hold the models.  So the drop-in is an overlay
(SURVEY.md §8b):

1. install compatibility shims for what the reference needs but a modern stack lacks
   (``np.int``/``np.float``; ``matplotlib``, ``webcolors``, ``prettytable`` stubs when absent;
   ``torchvision.models.utils.load_state_dict_from_url``; ``torchvision.models.resnet.model_urls``);
2. import the reference modules that define or re-export the hot-path names;
3. rebind those names - in every loaded reference module that holds them - to this package's
   implementations (same signatures);
4. ``runpy.run_path`` the driver as ``__main__``.

Nothing is copied from the reference; it executes from where it lies.
"""
from __future__ import annotations

import importlib
import os
import runpy
import sys
import types

#: hot-path name -> attribute of this package
REPLACED = (
    "get_max_preds", "accuracy",                                    # utils/keypoint_detection.py
    "find_keypoints_max", "compute_uv_from_heatmaps", "compute_uv_from_heatmaps2", "compute_uv_from_heatmaps3",
    "JointsMSELoss", "JointsKLLoss",                                # uda/model/loss.py
    "PseudoLabelGenerator", "PseudoLabelGenerator01", "PseudoLabelGenerator02", "PseudoLabelGenerator03",
    "RegressionDisparity", "RegressionDisparityx1", "RegressionDisparityx5", "RegressionDisparityx6",
    "RegressionDisparity2", "RegressionDisparity3", "RegressionDisparity4", "RegressionDisparity5",   # row f3
    "RegressionDisparity6", "RegressionDisparity7", "RegressionDisparity8",
    "RegressionDisparityx2", "RegressionDisparityx3", "RegressionDisparityx4", "JointsMSELoss0", "JointsKLLoss5",
)
#: rebound only with ``--device-targets``: the datasets call ``generate_target`` per sample from ``__getitem__``
#: inside DataLoader WORKER processes (train1.py defaults to ``--workers 4``, fork start method), where the parent's
#: CUDA context cannot be used - so by default the dataset side keeps the reference's numpy version and the batched
#: CUDA generator is an explicit choice (``generate_target_batch`` / ``target.DeviceTargetCollate``; needs
#: ``--workers 0`` when rebound per sample)
OPT_IN = ("generate_target",)                                       # uda/dataset/util.py

#: reference modules that define the names above (imported before rebinding)
DEFINING_MODULES = ("utils.keypoint_detection", "uda.model.loss", "uda.model.regda_4", "uda.model.regda_7",
                    "uda.dataset.util")


_BASIC_COLOURS = {"black": (0, 0, 0), "white": (255, 255, 255), "red": (255, 0, 0), "green": (0, 128, 0),
                  "blue": (0, 0, 255), "yellow": (255, 255, 0), "purple": (128, 0, 128), "orange": (255, 165, 0),
                  "cyan": (0, 255, 255), "magenta": (255, 0, 255), "pink": (255, 192, 203), "brown": (165, 42, 42),
                  "gray": (128, 128, 128), "grey": (128, 128, 128), "lime": (0, 255, 0), "navy": (0, 0, 128)}


def _stub_module(name):
    """Stand-in for an optional visualisation dependency that is not installed: any attribute is a
    no-op callable; ``webcolors.name_to_rgb`` (keypoint_dataset.py:52-55, drawing only) knows the basic names."""
    mod = types.ModuleType(name)
    mod.__dict__["__hp_stub__"] = True

    def _noop(*args, **kwargs):
        return None

    def _getattr(attr):
        if attr.startswith("__"):
            raise AttributeError(attr)
        if name == "webcolors" and attr == "name_to_rgb":
            return lambda colour, *a, **k: _BASIC_COLOURS.get(str(colour).lower(), (0, 0, 0))
        return _noop
    mod.__getattr__ = _getattr
    return mod


def install_shims():
    import numpy as np

    if not hasattr(np, "int"):
        np.int = int
    if not hasattr(np, "float"):
        np.float = float
    for name in ("matplotlib", "matplotlib.pyplot", "webcolors", "prettytable"):
        if name in sys.modules:
            continue
        try:
            importlib.import_module(name)
        except Exception:
            sys.modules[name] = _stub_module(name)
    if "matplotlib" in sys.modules and "matplotlib.pyplot" in sys.modules:
        setattr(sys.modules["matplotlib"], "pyplot", sys.modules["matplotlib.pyplot"])
    try:
        import torchvision.models.utils  # noqa: F401  (removed in torchvision >= 0.13)
    except Exception:
        try:
            import torch.hub
            import torchvision.models as tvm
            shim = types.ModuleType("torchvision.models.utils")
            shim.load_state_dict_from_url = torch.hub.load_state_dict_from_url
            sys.modules["torchvision.models.utils"] = shim
            tvm.utils = shim
            import torchvision.models.resnet as tvr
            if not hasattr(tvr, "model_urls"):
                tvr.model_urls = {k: "" for k in ("resnet18", "resnet34", "resnet50", "resnet101", "resnet152",
                                                  "resnext50_32x4d", "resnext101_32x8d", "wide_resnet50_2",
                                                  "wide_resnet101_2")}
        except Exception:
            pass


def install(ref_root: str, verbose: bool = False, device_targets: bool = False, route_upsample: bool = True,
            lazy_upsample: bool = True):
    """Shim, import and rebind.  Returns ``{module_name: [rebound names]}``.
    ``device_targets``: also rebind the per-sample ``generate_target`` (see OPT_IN).
    ``route_upsample``: send ``nn.Upsample(mode='bilinear')`` of detached CUDA fp32 heatmaps (train1.py:410-417) to
    ``hp_fuse_multiscale`` (row a12).  ``lazy_upsample``: the route returns ``fusion.LazyUpsample`` objects, so that the
    driver's own ``target5 = 0.5 * target + target1`` reaches ``RegressionDisparityx6`` as its two heads and is built inside
    the loss kernel (``--eager-upsample`` switches this off)."""
    ref_root = os.path.abspath(ref_root)
    if not os.path.isfile(os.path.join(ref_root, "utils", "keypoint_detection.py")):
        raise FileNotFoundError(f"{ref_root} does not look like the reference tree")
    sys.dont_write_bytecode = True
    install_shims()
    if ref_root not in sys.path:
        sys.path.insert(0, ref_root)
    pkg = importlib.import_module(__package__)
    originals = {}
    for modname in DEFINING_MODULES:
        mod = importlib.import_module(modname)
        for name in REPLACED + (OPT_IN if device_targets else ()):
            if hasattr(mod, name) and getattr(mod, name).__module__ == mod.__name__:
                originals.setdefault(name, []).append(getattr(mod, name))   # regda_4 and regda_7 both define some
    rebound = {}
    for modname, mod in list(sys.modules.items()):
        path = getattr(mod, "__file__", None) or ""
        if not path.startswith(ref_root):
            continue
        for name, origs in originals.items():
            if any(mod.__dict__.get(name) is o for o in origs):
                setattr(mod, name, getattr(pkg, name))
                rebound.setdefault(modname, []).append(name)
    if route_upsample:
        install_upsample_route(lazy=lazy_upsample)
    if verbose:
        for m in sorted(rebound):
            print(f"[hpb200 overlay] {m}: {', '.join(sorted(rebound[m]))}", file=sys.stderr)
    return rebound


_UPSAMPLE_ROUTED = False


_UPSAMPLE_LAZY = True


def install_upsample_route(lazy=True):
    """train1.py:410-417 / test.py:362-369 build ``nn.Upsample(size=64|32, mode='bilinear')`` inline and apply them
    to DETACHED adversarial heatmaps.  The drivers must stay unchanged, so the route is on ``nn.Upsample.forward``:
    a bilinear, align_corners-free upsample of a CUDA fp32 4-D tensor that carries no autograd history goes to the
    gather+blend kernel (``fusion.upsample_bilinear`` -> ``hp_fuse_multiscale``); everything else (other modes,
    tensors that need gradients - e.g. inside a model -, CPU tensors, scale_factor forms) takes torch's own path.

    ``lazy``: the route returns a ``fusion.LazyUpsample`` (the map is not built yet).  train1.py:424-428 only scales, adds and
    hands these maps to the disparity losses: ``0.5 * target + target1`` becomes ``fusion.FusedHeads``, which
    ``RegressionDisparityx6(..., mode='max')`` consumes as two heads (the fused map is interpolated inside the loss kernel and
    never written); ``target0`` is materialised by ``RegressionDisparityx5`` with one launch.  Any other use of such an object
    (a torch function, an attribute of a tensor, other arithmetic) materialises it first - same values as the eager route."""
    global _UPSAMPLE_ROUTED, _UPSAMPLE_LAZY
    _UPSAMPLE_LAZY = bool(lazy)
    if _UPSAMPLE_ROUTED:
        return
    import torch
    import torch.nn as nn
    fusion = importlib.import_module(__package__ + ".fusion")
    stock = nn.Upsample.forward

    def forward(self, input):
        size = self.size
        if (self.mode == "bilinear" and not self.align_corners and size is not None and isinstance(input, torch.Tensor)
                and input.is_cuda and input.dtype == torch.float32 and input.dim() == 4
                and not (input.requires_grad and torch.is_grad_enabled())):
            hw = (size, size) if isinstance(size, int) else tuple(int(v) for v in size)
            if len(hw) == 2 and hw[0] >= input.shape[2] and hw[1] >= input.shape[3]:
                return fusion.LazyUpsample(input, hw) if _UPSAMPLE_LAZY else fusion.upsample_bilinear(input, hw)
        return stock(self, input)

    nn.Upsample.forward = forward
    _UPSAMPLE_ROUTED = True


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    ref = os.environ.get("HP_REF_DIR")
    verbose = device_targets = False
    lazy_upsample = True
    plugins = []
    while argv and argv[0].startswith("--"):
        if argv[0] == "--ref" and len(argv) > 1:
            ref = argv[1]
            argv = argv[2:]
        elif argv[0] == "--plugin" and len(argv) > 1:
            plugins.append(argv[1])
            argv = argv[2:]
        elif argv[0] == "--device-targets":
            device_targets = True
            argv = argv[1:]
        elif argv[0] == "--verbose":
            verbose = True
            argv = argv[1:]
        elif argv[0] == "--eager-upsample":
            lazy_upsample = False
            argv = argv[1:]
        else:
            break
    if not argv or ref is None:
        print(__doc__)
        print("error: give the reference tree with --ref DIR or HP_REF_DIR", file=sys.stderr)
        return 2
    script = argv[0] if os.path.isabs(argv[0]) else os.path.join(ref, argv[0])
    install(ref, verbose=verbose, device_targets=device_targets, lazy_upsample=lazy_upsample)
    for path in plugins:       # launcher-side registrations (e.g. a synthetic dataset / a stand-in backbone by name)
        runpy.run_path(path, run_name="__hpb200_plugin__")
    sys.argv = [script] + argv[1:]
    runpy.run_path(script, run_name="__main__")
    return 0
