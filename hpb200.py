"""Short import alias: ``import hpb200`` == the package directory
``domain-adaptative-hand-pose-estimation_b200/`` (whose name is not a Python identifier)."""
import importlib
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
if _here not in sys.path:
    sys.path.insert(0, _here)
_pkg = importlib.import_module("domain-adaptative-hand-pose-estimation_b200")
sys.modules[__name__] = _pkg

if __name__ == "__main__":          # `python hpb200.py [--ref DIR] train1.py ...` : overlay launcher
    sys.exit(importlib.import_module(_pkg.__name__ + ".overlay").main())
