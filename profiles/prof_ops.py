#!/usr/bin/env python
"""A few calls of the secondary operators for an ncu capture:
    ncu --set full --clock-control none --import-source on -k regex:"fuse_staged|regdisp" -c 8 -o prof python profiles/prof_ops.py"""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
hp = importlib.import_module("domain-adaptative-hand-pose-estimation_b200")
dev = torch.device("cuda", 0)
K = 21
with torch.no_grad():
    hi = hp.synth.make_device_batch(40, 256, K, 128, 128, image_size=512, device=dev)["pred"]
    mid = torch.nn.functional.avg_pool2d(hi, 2)
    lo = torch.nn.functional.avg_pool2d(hi, 4)
    tgt = torch.randint(0, 128, (256, K, 2), device=dev).float()
    ev = hp.MultiscaleEval(K)
    for _ in range(2):
        ev(lo, mid, hi, tgt)
    del hi, mid, lo
    y = hp.synth.make_device_batch(20, 512, K, 64, 64, device=dev)["pred"]
    adv = hp.synth.make_device_batch(30, 512, K, 64, 64, device=dev)["pred"]
    t5 = hp.fuse_multiscale(torch.nn.functional.avg_pool2d(adv, 4), torch.nn.functional.avg_pool2d(adv, 2), 64, 32)[0]
    rd6 = hp.RegressionDisparityx6(hp.PseudoLabelGenerator(K, 64, 64), hp.JointsKLLoss(epsilon=1e-7))
    for _ in range(2):
        rd6(y, adv, None, None, "min")
        rd6(y, adv, t5, None, "max")
torch.cuda.synchronize()
print("ok")
