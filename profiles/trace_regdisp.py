#!/usr/bin/env python
"""Timeline of the dense 'max' disparity kernel (csrc/hp_regdisp_dense.cuh) from its own globaltimer stamps
(hp_debug_regdisp_trace).

    python profiles/trace_regdisp.py [--fused] [--batch 512] [--out gpurun_out/rd_trace.json]

Per consumer warp and map: wait (data landed - wait begins; without a fused map this includes the softmax-maximum pass),
lp (per-sample label seen - data landed), proc (map done - label seen); per sample the builder's begin / published;
per block entry / exit.  Times in ns relative to the earliest block entry."""
import argparse, importlib, json, os, sys
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
hp = importlib.import_module("domain-adaptative-hand-pose-estimation_b200")
L = importlib.import_module("domain-adaptative-hand-pose-estimation_b200._lib")

ap = argparse.ArgumentParser()
ap.add_argument("--fused", action="store_true")
ap.add_argument("--heads", action="store_true", help="with --fused: leave the map unfused (hp.FusedHeads, built inside the kernel)")
ap.add_argument("--batch", type=int, default=512)
ap.add_argument("--out", default=None)
args = ap.parse_args()

dev = torch.device("cuda", 0)
lib = L.load()
K, B = 21, args.batch
y = hp.synth.make_device_batch(20, B, K, 64, 64, device=dev)["pred"]
advs = [hp.synth.make_device_batch(30 + i, B, K, 64, 64, device=dev)["pred"] for i in range(3)]
f = None
if args.fused:
    a32 = torch.nn.functional.avg_pool2d(advs[0], 2)
    a16 = torch.nn.functional.avg_pool2d(advs[0], 4)
    f = hp.FusedHeads(a16, a32) if args.heads else hp.fuse_multiscale(a16, a32, 64, 32)[0]
rd6 = hp.RegressionDisparityx6(hp.PseudoLabelGenerator(K, 64, 64), hp.JointsKLLoss(epsilon=1e-7))
words = int(lib.hp_debug_regdisp_trace_words())
buf = torch.zeros(words, dtype=torch.int64, device=dev)
with torch.no_grad():
    for i in range(6):
        rd6(y, advs[i % 3], f, None, "max")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(20):
        rd6(y, advs[i % 3], f, None, "max")
    e1.record()
    torch.cuda.synchronize()
    us_step = 1e3 * e0.elapsed_time(e1) / 20
    L.call("hp_debug_regdisp_trace", L.ptr(buf), words)
    rd6(y, advs[0], f, None, "max")
    torch.cuda.synchronize()
    L.call("hp_debug_regdisp_trace", None, 0)

NM, NWMAX = 16, 16
BW = 8 + NWMAX * NM * 8 + 16
t = buf.cpu().numpy().astype(np.int64).reshape(-1, BW)
t = t[t[:, 0] != 0]
t0 = t[:, 0].min()
entry, exit_ = t[:, 0] - t0, t[:, 1] - t0
maps = t[:, 8:8 + NWMAX * NM * 8].reshape(-1, NWMAX, NM, 8)
bld = t[:, 8 + NWMAX * NM * 8:].reshape(-1, 8, 2)
used = maps[..., 6] != 0
rel = np.where(maps != 0, maps - t0, 0)
wait = (rel[..., 1] - rel[..., 0])[used]
lpw = (rel[..., 2] - rel[..., 1])[used]
proc = (rel[..., 6] - rel[..., 2])[used]
first = used & (np.arange(NM)[None, None, :] == 0)
later = used & (np.arange(NM)[None, None, :] > 0)


def q(x):
    x = np.asarray(x)
    return {"p10": float(np.percentile(x, 10)), "p50": float(np.percentile(x, 50)), "p90": float(np.percentile(x, 90)),
            "max": float(x.max()), "mean": float(x.mean())} if x.size else {}


bu = bld[..., 1] != 0
res = {
    "us_per_step_decode_plus_loss": us_step, "fused": args.fused, "heads": args.heads, "batch": B, "blocks": int(t.shape[0]),
    "block_entry_ns": q(entry), "block_exit_ns": q(exit_), "kernel_span_ns": float(exit_.max()),
    "first_lp_published_ns": q((bld[:, 0, 1] - t0)), "builder_build_ns": q((bld[..., 1] - bld[..., 0])[bu]),
    "builder_begin_by_sample_ns": [q((bld[:, r, 0] - t0)[bld[:, r, 1] != 0]) for r in range(6)],
    "map_wait_ns": q(wait), "map_lp_wait_ns": q(lpw), "map_proc_ns": q(proc),
    "first_map": {"wait": q((rel[..., 1] - rel[..., 0])[first]), "lp": q((rel[..., 2] - rel[..., 1])[first]),
                  "done_at": q(rel[..., 6][first])},
    "later_maps": {"wait": q((rel[..., 1] - rel[..., 0])[later]), "lp": q((rel[..., 2] - rel[..., 1])[later]),
                   "proc": q((rel[..., 6] - rel[..., 2])[later]),
                   "patch": q((rel[..., 3] - rel[..., 2])[later]), "passes": q((rel[..., 4] - rel[..., 3])[later]),
                   "reduce_request": q((rel[..., 5] - rel[..., 4])[later]), "closure": q((rel[..., 6] - rel[..., 5])[later])},
    "maps_done_at_by_index_ns": [q(rel[..., j, 6][used[..., j]]) for j in range(NM) if used[..., j].any()],
    "maps_per_warp": q(used.sum(axis=2).ravel()),
    "prologue_warp0_ns": {"slots_ready": q(t[:, 5] - t0), "joints_done": q(t[:, 6] - t0), "clip_done": q(t[:, 7] - t0),
                          "barrier_passed": q(t[:, 2] - t0)},
}
print(json.dumps(res, indent=1))
if args.out:
    with open(args.out, "w") as fo:
        json.dump(res, fo, indent=1)
