#!/usr/bin/env python
"""End-to-end host path (HeatmapPipeline.run_host, configs[1]: 256 x 21 x 64 x 64 from pinned host memory) against the slab size,
next to the raw pinned host->device copy rate of the same 88.1 MB (one cudaMemcpyAsync; CUDA events)."""
import importlib, json, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
hp = importlib.import_module("domain-adaptative-hand-pose-estimation_b200")
dev = torch.device("cuda", 0)
B, K, S = 256, 21, 64
hs = hp.synth.make_host_batch(77, B, K, S, S)
pred = torch.from_numpy(hs["pred"]).pin_memory()
joints = torch.from_numpy(hs["joints"]).pin_memory()
vis = torch.from_numpy(hs["vis"]).pin_memory()
dst = torch.empty_like(pred, device=dev)
out = {}
for _ in range(3):
    dst.copy_(pred, non_blocking=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    dst.copy_(pred, non_blocking=True)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
out["raw_h2d_GBps"] = pred.numel() * 4 / (ms * 1e-3) / 1e9
out["raw_h2d_ms"] = ms
pipe = hp.HeatmapPipeline(K, (S, S), (4 * S, 4 * S), sigma=2, losses=("mse", "kl"))
for slab in (16, 32, 64, 128, 256):
    for _ in range(3):
        pipe.run_host(pred, joints, vis, slab=slab, want_pred_xy=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 20
    for _ in range(n):
        pipe.run_host(pred, joints, vis, slab=slab, want_pred_xy=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / n
    out[f"slab{slab}_ms"] = dt * 1e3
    out[f"slab{slab}_Mmaps_s"] = B * K / dt / 1e6
print(json.dumps(out, indent=1))
