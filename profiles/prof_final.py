#!/usr/bin/env python
"""The kernels that changed last in round 1 (staged loss / accuracy at 64x64, fusion block kernel) for an ncu capture:
    ncu --set full --clock-control none --import-source on -k regex:"staged|fuse_block" -c 5 -o gpurun_out/final python profiles/prof_final.py"""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
hp = importlib.import_module("domain-adaptative-hand-pose-estimation_b200")
dev = torch.device("cuda", 0)
K, B, S = 21, 256, 64
s = hp.synth.make_device_batch(10, B, K, S, S, image_size=256, device=dev)
tg = hp.generate_target_batch(s["joints"], s["vis"], (S, S), 2, (256, 256))
hi = hp.synth.make_device_batch(40, B, K, 128, 128, image_size=512, device=dev)["pred"]
mid, lo = torch.nn.functional.avg_pool2d(hi, 2), torch.nn.functional.avg_pool2d(hi, 4)
tgt = torch.randint(0, 128, (B, K, 2), device=dev).float()
a16, a32 = torch.nn.functional.avg_pool2d(s["pred"], 4), torch.nn.functional.avg_pool2d(s["pred"], 2)
mse, kl, ev = hp.JointsMSELoss(), hp.JointsKLLoss(epsilon=1e-7), hp.MultiscaleEval(K)
with torch.no_grad():
    kl(s["pred"], tg[0], tg[1])
    mse(s["pred"], tg[0], tg[1])
    hp.pck(s["pred"], tg[0])
    ev(lo, mid, hi, tgt)
    hp.fuse_multiscale(a16, a32, 64, 32)
torch.cuda.synchronize()
print("ok")
