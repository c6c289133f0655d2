#!/usr/bin/env python
"""Make the reference travel to the GPU box (TEST INFRASTRUCTURE; run by oracle/Makefile from __graft_entry__.build()).

The reference is pure Python, so "building" it means copying its *.py files - nothing else - from /root/reference
(read-only, present only in the build container) into oracle/_ref/reference/.  That directory is git-ignored (it never
enters the history: reference sources are not part of this repo) but not gpurun-ignored, so it is on the GPU box next
to the built libhp_b200.so.  There it serves
  * `bench.py --impl reference` / `cpu_baseline` (kind "reference": the reference's OWN functions are what is timed),
  * tests/test_reference_on_gpu_box.py (oracle == reference where the tree exists only as this copy),
  * tests/test_train1_overlay.py (row f1: the unchanged train1.py through the overlay on a GPU).
Nothing in the product package reads it."""
import os
import shutil
import sys

SRC = os.environ.get("HP_REF_SRC", "/root/reference")
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "reference")


def main():
    if not os.path.isfile(os.path.join(SRC, "utils", "keypoint_detection.py")):
        print(f"ship_reference: {SRC} not present - keeping whatever oracle/_ref/reference already holds")
        return 0
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    n = 0
    for root, dirs, files in os.walk(SRC):
        dirs[:] = [d for d in dirs if d not in ("__pycache__", ".git")]
        for f in files:
            if not f.endswith(".py"):
                continue
            rel = os.path.relpath(os.path.join(root, f), SRC)
            out = os.path.join(DST, rel)
            os.makedirs(os.path.dirname(out), exist_ok=True)
            shutil.copyfile(os.path.join(root, f), out)
            n += 1
    print(f"ship_reference: {n} .py files -> {DST}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
