"""Freeze golden vectors FROM THE REAL REFERENCE  (run here, where /root/reference exists).

    python -m oracle.gen_golden            # writes tests/golden/<case>.npz + MANIFEST.json
    python -m oracle.gen_golden step_b     # only the named cases (a new case is added without touching the other files)

Evaluates every case of ``oracle/cases.py`` on the reference's own code (loaded in place by
``oracle/ref_loader.py``) and stores the outputs.  Inputs are regenerated from seeds by the
tests.  Arrays above ``DIGEST_ABOVE`` elements are stored as a digest (strided sample + float64
L1 + L2 norms) to keep the fixtures small; ``digest()`` is applied to the candidate side too.
"""
from __future__ import annotations

import hashlib
import json
import os
import sys
import warnings

import numpy as np

DIGEST_ABOVE = 1 << 15
DIGEST_STRIDE = 29          # co-prime with every map width in use

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN_DIR = os.path.join(_ROOT, "tests", "golden")


def digest(outputs: dict) -> dict:
    out = {}
    for key, a in outputs.items():
        a = np.asarray(a)
        if a.size <= DIGEST_ABOVE:
            out[key] = a
            continue
        flat = a.reshape(-1)
        out[key + "#sample"] = flat[::DIGEST_STRIDE].copy()
        f64 = flat.astype(np.float64)
        out["c:" + key[2:] + "#l1"] = np.asarray([np.abs(f64).sum()])   # signed sums of gradients cancel to ~0
        out["c:" + key[2:] + "#l2"] = np.asarray([np.sqrt((f64 * f64).sum())])
        out["x:" + key[2:] + "#shape"] = np.asarray(a.shape, dtype=np.int64)
    return out


def main():
    warnings.filterwarnings("ignore", category=SyntaxWarning)
    if _ROOT not in sys.path:
        sys.path.insert(0, _ROOT)
    from oracle import cases, ref_loader
    import torch

    torch.manual_seed(0)
    ns = ref_loader.load()
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    manifest = {"source": "real reference executed in place from /root/reference",
                "numpy": np.__version__, "torch": torch.__version__, "files": {}}
    only = set(sys.argv[1:])     # python -m oracle.gen_golden [case ...]: freeze only these (the other files and their entries stay)
    if only:
        with open(os.path.join(GOLDEN_DIR, "MANIFEST.json")) as f:
            manifest["files"] = json.load(f)["files"]
    for name, fn in cases.CASES.items():
        if only and name not in only:
            continue
        out = digest(fn(ns, "cpu"))
        path = os.path.join(GOLDEN_DIR, f"{name}.npz")
        np.savez_compressed(path, **out)
        with open(path, "rb") as f:
            sha = hashlib.sha256(f.read()).hexdigest()
        manifest["files"][f"{name}.npz"] = {"keys": len(out), "bytes": os.path.getsize(path), "sha256": sha}
        print(f"{name:14s} {len(out):4d} arrays  {os.path.getsize(path) / 1024:8.1f} KiB")
    with open(os.path.join(GOLDEN_DIR, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
