#!/usr/bin/env python
"""A few loss / accuracy / target calls for an ncu capture:
    ncu --set full --clock-control none --import-source on -k regex:"loss_fwd|accuracy_kernel|gaussian_target|soft_argmax" -c 6 -o gpurun_out/losses python profiles/prof_losses.py"""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
hp = importlib.import_module("domain-adaptative-hand-pose-estimation_b200")
dev = torch.device("cuda", 0)
K, B, S = 21, 256, 64
s = hp.synth.make_device_batch(10, B, K, S, S, image_size=256, device=dev)
tg = hp.generate_target_batch(s["joints"], s["vis"], (S, S), 2, (256, 256))
mse, kl = hp.JointsMSELoss(), hp.JointsKLLoss(epsilon=1e-7)
with torch.no_grad():
    for _ in range(2):
        kl(s["pred"], tg[0], tg[1])
        mse(s["pred"], tg[0], tg[1])
        hp.pck(s["pred"], tg[0])
        hp.compute_uv_from_heatmaps3(s["pred"])
torch.cuda.synchronize()
print("ok")
