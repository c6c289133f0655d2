#!/usr/bin/env python
"""BASELINE.json configs[3]: multiscale 32/64/128 fusion + decode + PCK, batch 2048 sharded over the GPUs of one box
(256 samples per GPU), one NCCL all-reduce of the PCK counts per step.  Launch like bench.py:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
           profiles/bench_c4.py --steps 300 --warmup 20          (N = 1: plain `python profiles/bench_c4.py`)

Timing: CUDA events on the launching stream around the K steps, barrier + synchronize on both sides, MAX over ranks;
3 rotating input sets of 462 MB per GPU (> the 126 MB L2).  Rank 0 prints one JSON line (heatmaps/s = fused 128x128 maps)."""
import argparse, importlib, json, os, sys
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
hp = importlib.import_module("domain-adaptative-hand-pose-estimation_b200")

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=300)
ap.add_argument("--warmup", type=int, default=20)
ap.add_argument("--per-gpu-batch", type=int, default=256)
args = ap.parse_args()
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
K, B = 21, args.per_gpu_batch
with torch.no_grad():
    hi = [hp.synth.make_device_batch(4000 + 97 * rank + i, B, K, 128, 128, image_size=512, device=dev)["pred"] for i in range(3)]
    mid = [torch.nn.functional.avg_pool2d(h, 2) for h in hi]
    lo = [torch.nn.functional.avg_pool2d(h, 4) for h in hi]
    tgt = [torch.randint(0, 128, (B, K, 2), device=dev).float() for _ in range(3)]
    ev = hp.MultiscaleEval(K)
    for i in range(args.warmup):
        ev(lo[i % 3], mid[i % 3], hi[i % 3], tgt[i % 3])
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        acc, _, _ = ev(lo[i % 3], mid[i % 3], hi[i % 3], tgt[i % 3])
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
ms_step = ms.item() / args.steps
maps = world * B * K
bytes_map = 4096 + 16384 + 65536 + 16
peak = 6450.3
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
if rank == 0:
    gbs = B * K * bytes_map / (ms_step * 1e-3) / 1e9
    print(json.dumps({"metric": "heatmaps/sec (fuse 32/64/128 + decode + PCK, sharded, PCK all-reduce)", "value": maps / (ms_step * 1e-3),
                      "unit": "heatmaps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
                      "scaling": "weak", "per_gpu_batch": B, "per_gpu_GBps": gbs, "frac_of_measured_hbm_per_gpu": gbs / peak,
                      "avg_acc": float(acc[K].item()), "collective": "nccl all_reduce of 2K counts" if world > 1 else "none"}))
if world > 1:
    dist.destroy_process_group()
