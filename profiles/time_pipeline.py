#!/usr/bin/env python
"""Quick A/B timer for the fused pipeline kernel under the library's experiment switches (HP_PIPE_SHAPE, HP_PIPE_EPILOGUE,
...: read once per process, so run one process per variant).

    HP_PIPE_SHAPE=5 python profiles/time_pipeline.py [--batch 256] [--side 64] [--launches 2000] [--tag name]

Prints one JSON line: us per launch in a train (overlap=True), serialised back-to-back (overlap=False), and isolated
(one launch between two events on an idle GPU), plus the result scalars (to eyeball parity between variants)."""
import argparse, importlib, json, os, statistics, sys
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
hp = importlib.import_module("domain-adaptative-hand-pose-estimation_b200")

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--side", type=int, default=64)
ap.add_argument("--launches", type=int, default=2000)
ap.add_argument("--tag", default="")
args = ap.parse_args()
dev = torch.device("cuda", 0)
S, B = args.side, args.batch
n_sets = 8 if S <= 64 else 3
sets = [hp.synth.make_device_batch(100 + i, B, 21, S, S, image_size=4 * S, device=dev) for i in range(n_sets)]
pipe = hp.HeatmapPipeline(heatmap_size=(S, S), image_size=(4 * S, 4 * S), kl_epsilon=1e-7, device=dev)
outs = [pipe.alloc_outputs(B, dev) for _ in range(8)]


def bench(overlap, n):
    ls = [pipe.plan(sets[i % n_sets]["pred"], sets[i % n_sets]["joints"], sets[i % n_sets]["vis"], out=outs[i], overlap=overlap)[0]
          for i in range(8)]
    for i in range(16):
        ls[i % 8]()
    torch.cuda.synchronize()
    vals = []
    for rep in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            ls[i % 8]()
        e1.record()
        torch.cuda.synchronize()
        vals.append(1e3 * e0.elapsed_time(e1) / n)
    return ls, vals


ls_t, train = bench(True, args.launches)
ls_s, serial = bench(False, max(200, args.launches // 4))
serial40 = []
for rep in range(15):      # short serialised bursts: is the long-burst figure a clock / power effect?
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(40):
        ls_s[i % 8]()
    e1.record()
    torch.cuda.synchronize()
    serial40.append(1e3 * e0.elapsed_time(e1) / 40)
short = []
for rep in range(9):       # the driver's shape: 20-launch trains
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(20):
        ls_t[i % 8]()
    e1.record()
    torch.cuda.synchronize()
    short.append(1e3 * e0.elapsed_time(e1) / 20)
iso = []
for i in range(30):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ls_s[i % 8]()
    e1.record()
    torch.cuda.synchronize()
    iso.append(1e3 * e0.elapsed_time(e1))
r = outs[0].host()
alg = (S * S * 4 + 32) * B * 21
f = lambda us: alg / (us * 1e-6) / 1e9 / 6450.3
print(json.dumps({"tag": args.tag, "env": {k: v for k, v in os.environ.items() if k.startswith("HP_")},
                  "train_us": round(statistics.median(train), 3), "train20_us": round(statistics.median(short), 3),
                  "serial_us": round(statistics.median(serial), 3), "serial_all": [round(v, 2) for v in serial],
                  "serial40_us": round(statistics.median(serial40), 3), "isolated_us": round(statistics.median(iso), 3),
                  "frac_train": round(f(statistics.median(train)), 3), "frac_serial": round(f(statistics.median(serial)), 3),
                  "frac_isolated": round(f(statistics.median(iso)), 3), "mse": r["mse"], "kl": r["kl"], "avg_acc": r["avg_acc"]}))
