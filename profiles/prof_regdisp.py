#!/usr/bin/env python
"""The regression-disparity operators (C3: 512x21x64x64) for an ncu capture / launch list:
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/rd_launches.csv python profiles/prof_regdisp.py
    ncu --set full --clock-control none --import-source on -k regex:regdisp_staged -c 6 -o gpurun_out/rd python profiles/prof_regdisp.py"""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
hp = importlib.import_module("domain-adaptative-hand-pose-estimation_b200")
dev = torch.device("cuda", 0)
K, B = 21, 512
y = hp.synth.make_device_batch(20, B, K, 64, 64, device=dev)["pred"]
adv = hp.synth.make_device_batch(30, B, K, 64, 64, device=dev)["pred"]
a32 = torch.nn.functional.avg_pool2d(adv, 2)
a16 = torch.nn.functional.avg_pool2d(adv, 4)
t5, t0 = hp.fuse_multiscale(a16, a32, 64, 32)
kl = hp.JointsKLLoss(epsilon=1e-7)
rd6 = hp.RegressionDisparityx6(hp.PseudoLabelGenerator(K, 64, 64), kl)
rd5 = hp.RegressionDisparityx5(hp.PseudoLabelGenerator03(K), kl)
rd1 = hp.RegressionDisparityx1(hp.PseudoLabelGenerator01(K), kl)
reps = int(os.environ.get("REPS", "2"))
for _ in range(reps):
    for args, rd in (((y, adv.clone().requires_grad_(True), None, None, "min"), rd6),
                     ((y, adv.clone().requires_grad_(True), None, None, "max"), rd6),
                     ((y, adv.clone().requires_grad_(True), t5, None, "max"), rd6),
                     ((y, adv.clone().requires_grad_(True), hp.FusedHeads(a16, a32), None, "max"), rd6),
                     ((y, a32.clone().requires_grad_(True), None, None, "min"), rd5),
                     ((y, a32.clone().requires_grad_(True), t0, None, "max"), rd5),
                     ((y, a16.clone().requires_grad_(True), None, "min"), rd1),
                     ((y, a16.clone().requires_grad_(True), None, "max"), rd1)):
        rd(*args).backward()
torch.cuda.synchronize()
print("ok")
