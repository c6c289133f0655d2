import importlib, os, sys, cProfile, pstats, time
import torch
sys.path.insert(0, os.getcwd())
hp = importlib.import_module("domain-adaptative-hand-pose-estimation_b200")
dev = torch.device("cuda", 0)
B, K, S = 256, 21, 64
s = hp.synth.make_device_batch(10, B, K, S, S, image_size=256, device=dev)
tg = hp.generate_target_batch(s["joints"], s["vis"], (S, S), 2, (256, 256))
mse = hp.JointsMSELoss()
pred = s["pred"].clone().requires_grad_(True)
def step():
    pred.grad = None
    mse(pred, tg[0], tg[1]).backward()
def fwd():
    with torch.no_grad():
        mse(pred, tg[0], tg[1])
for f in (fwd, step):
    for _ in range(200): f()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(2000): f()
    host = (time.perf_counter() - t) / 2000 * 1e6
    torch.cuda.synchronize()
    print(f.__name__, "host us/iter", host)
    pr = cProfile.Profile(); pr.enable()
    for _ in range(2000): f()
    pr.disable(); torch.cuda.synchronize()
    pstats.Stats(pr).sort_stats("tottime").print_stats(14)
