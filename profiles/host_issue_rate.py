#!/usr/bin/env python
"""Host-side cost of issuing one step (Python + ctypes + cudaLaunchKernelEx), against the device time per step.
    python profiles/host_issue_rate.py          (single process, configs[1])"""
import importlib, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
hp = importlib.import_module("domain-adaptative-hand-pose-estimation_b200")
dev = torch.device("cuda", 0)
sets = [hp.synth.make_device_batch(100 + i, 256, device=dev) for i in range(8)]
pipe = hp.HeatmapPipeline(kl_epsilon=1e-7, device=dev)
outs = [pipe.alloc_outputs(256, dev) for _ in range(8)]
plans = [pipe.plan(s["pred"], s["joints"], s["vis"], o, True, True)[0] for s, o in zip(sets, outs)]
for mode in ("pipe()", "plan.launch()"):
    for n in (200, 800):
        for i in range(50):
            pipe(sets[i % 8]["pred"], sets[i % 8]["joints"], sets[i % 8]["vis"], out=outs[i % 8], overlap=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        if mode == "pipe()":
            for i in range(n):
                s = sets[i % 8]
                pipe(s["pred"], s["joints"], s["vis"], out=outs[i % 8], overlap=True)
        else:
            for i in range(n):
                plans[i % 8]()
        e1.record()
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        print(f"{mode:14s} n={n:4d}: host issue {1e6 * (t1 - t0) / n:6.2f} us/step ; device {1e3 * e0.elapsed_time(e1) / n:6.2f} us/step ; "
              f"wall to completion {1e6 * (t2 - t0) / n:6.2f} us/step")
