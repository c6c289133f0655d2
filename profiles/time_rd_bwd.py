#!/usr/bin/env python
"""Forward and forward+backward times of the disparity operators at 512 x 21 (CUDA events, 30 reps, 3 input sets)."""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
hp = importlib.import_module("domain-adaptative-hand-pose-estimation_b200")
dev = torch.device("cuda", 0)
K, B = 21, 512
ys = [hp.synth.make_device_batch(20 + i, B, K, 64, 64, device=dev)["pred"] for i in range(3)]
advs = [hp.synth.make_device_batch(30 + i, B, K, 64, 64, device=dev)["pred"] for i in range(3)]
a32 = [torch.nn.functional.avg_pool2d(a, 2) for a in advs]
a16 = [torch.nn.functional.avg_pool2d(a, 4) for a in advs]
t5, t0 = zip(*[hp.fuse_multiscale(a16[i], a32[i], 64, 32) for i in range(3)])
kl = hp.JointsKLLoss(epsilon=1e-7)
rd6 = hp.RegressionDisparityx6(hp.PseudoLabelGenerator(K, 64, 64), kl)
rd5 = hp.RegressionDisparityx5(hp.PseudoLabelGenerator03(K), kl)
rd1 = hp.RegressionDisparityx1(hp.PseudoLabelGenerator01(K), kl)
cases = {
    "x6 min": lambda i, a: rd6(ys[i], a, None, None, "min"), "x6 max": lambda i, a: rd6(ys[i], a, None, None, "max"),
    "x6 max+t5": lambda i, a: rd6(ys[i], a, t5[i], None, "max"),
    "x5 min": lambda i, a: rd5(ys[i], a, None, None, "min"), "x5 max+t0": lambda i, a: rd5(ys[i], a, t0[i], None, "max"),
    "x1 min": lambda i, a: rd1(ys[i], a, None, "min"), "x1 max": lambda i, a: rd1(ys[i], a, None, "max"),
}
src = {"x6": advs, "x5": a32, "x1": a16}
reps = 30
for name, fn in cases.items():
    heads = [t.clone().requires_grad_(True) for t in src[name[:2]]]
    res = {}
    for what in ("fwd", "fwd+bwd"):
        def step(i):
            if what == "fwd":
                with torch.no_grad():
                    fn(i % 3, heads[i % 3])
            else:
                heads[i % 3].grad = None
                fn(i % 3, heads[i % 3]).backward()
        for i in range(5):
            step(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(reps):
            step(i)
        e1.record()
        torch.cuda.synchronize()
        res[what] = 1e3 * e0.elapsed_time(e1) / reps
    print(f"{name:12s} fwd {res['fwd']:7.1f} us   fwd+bwd {res['fwd+bwd']:7.1f} us   (bwd ~ {res['fwd+bwd'] - res['fwd']:6.1f} us)")
