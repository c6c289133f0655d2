#!/usr/bin/env python
"""Timeline of the TMA-staged pipeline kernel from its own clock stamps (hp_debug_pipeline_trace).

    python profiles/trace_pipeline.py [--overlap] [--launches N] [--out gpurun_out/trace.json]

Runs a train of back-to-back launches of configs[1] (256x21x64x64), keeps the stamps of the last two
launches and prints: block entry/exit spread (globaltimer), and per map the four phases
  wait  = data landed - wait begins      (the warp had nothing to do)
  load  = data landed - refill issued of the previous map of the same warp (bulk-copy latency)
  proc  = refill issued - data landed    (both passes over shared memory + patch + reductions)
  close = map closed - refill issued     (float64 closure, stores)
in nanoseconds at the SM clock reported by the run."""
import argparse, importlib, json, os, sys
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
hp = importlib.import_module("domain-adaptative-hand-pose-estimation_b200")
L = importlib.import_module("domain-adaptative-hand-pose-estimation_b200._lib")

ap = argparse.ArgumentParser()
ap.add_argument("--overlap", action="store_true")
ap.add_argument("--launches", type=int, default=40)
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--out", default=None)
args = ap.parse_args()

dev = torch.device("cuda", 0)
lib = L.load()
words = int(lib.hp_debug_pipeline_trace_words())
buf = torch.zeros(2 * words, dtype=torch.int64, device=dev)
sets = [hp.synth.make_device_batch(100 + i, args.batch, device=dev) for i in range(8)]
pipe = hp.HeatmapPipeline(kl_epsilon=1e-7, device=dev)
outs = [pipe.alloc_outputs(args.batch, dev) for _ in range(8)]
for i in range(16):
    pipe(sets[i % 8]["pred"], sets[i % 8]["joints"], sets[i % 8]["vis"], out=outs[i % 8], overlap=args.overlap)
torch.cuda.synchronize()
L.call("hp_debug_pipeline_trace", L.ptr(buf), words * 2)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(args.launches):
    pipe(sets[i % 8]["pred"], sets[i % 8]["joints"], sets[i % 8]["vis"], out=outs[i % 8], overlap=args.overlap)
e1.record()
torch.cuda.synchronize()
L.call("hp_debug_pipeline_trace", None, 0)
us_per_launch = 1e3 * e0.elapsed_time(e1) / args.launches
t = buf.cpu().numpy().astype(np.uint64).reshape(2, -1)
HDR = 8
BW = HDR + 16 * 8 * 4
n_blocks = words // BW
last, prev = (args.launches - 1) & 1, (args.launches - 2) & 1

def unpack(slot):
    b = t[slot].reshape(n_blocks, BW)
    used = b[:, 0] != 0                      # blocks that ran (grid <= n_blocks)
    b = b[used]
    hdr = b[:, :HDR].astype(np.int64)
    maps = b[:, HDR:].reshape(-1, 16, 8, 4).astype(np.int64)
    return hdr, maps

hdr, maps = unpack(last)
hdr_p, _ = unpack(prev)
t0 = hdr[:, 0].min()
# SM clock from block lifetime: cycles / ns
life_ns = (hdr[:, 2] - hdr[:, 0]).astype(np.float64)
life_cy = (hdr[:, 3] - hdr[:, 1]).astype(np.float64)
ghz = float(np.median(life_cy / np.maximum(life_ns, 1)))
rep = {"overlap": bool(args.overlap), "us_per_launch": us_per_launch, "sm_ghz_est": ghz,
       "block_entry_ns": {"min": 0, "median": float(np.median(hdr[:, 0] - t0)), "max": int((hdr[:, 0] - t0).max())},
       "block_exit_ns": {"min": int((hdr[:, 2] - t0).min()), "median": float(np.median(hdr[:, 2] - t0)), "max": int((hdr[:, 2] - t0).max())},
       "prev_launch_exit_ns_rel": {"min": int((hdr_p[:, 2] - t0).min()), "max": int((hdr_p[:, 2] - t0).max())},
       "prev_launch_entry_ns_rel": {"min": int((hdr_p[:, 0] - t0).min()), "max": int((hdr_p[:, 0] - t0).max())}}
valid = maps[..., 3] != 0
def stats(x):
    x = np.asarray(x, dtype=np.float64) / ghz
    return {"n": int(x.size), "p10": float(np.percentile(x, 10)), "median": float(np.median(x)), "p90": float(np.percentile(x, 90)), "max": float(x.max())} if x.size else {}
wait = (maps[..., 1] - maps[..., 0])[valid]
proc = (maps[..., 2] - maps[..., 1])[valid]
close = (maps[..., 3] - maps[..., 2])[valid]
rep["wait_ns"] = stats(wait); rep["proc_ns"] = stats(proc); rep["close_ns"] = stats(close)
v2 = valid[:, :, 1:] & valid[:, :, :-1]
load = (maps[:, :, 1:, 1] - maps[:, :, :-1, 2])[v2]
rep["load_ns_refill_to_landed"] = stats(load)
first_ready = (maps[:, :, 0, 1] - hdr[:, 1][:, None])[valid[:, :, 0]]
rep["first_data_after_entry_ns"] = stats(first_ready)
last_close = np.where(valid, maps[..., 3], 0).max(axis=(1, 2)) 
rep["exit_after_last_close_ns"] = stats(hdr[:, 3] - last_close)
rep["barrier_after_last_close_ns"] = stats(hdr[:, 5] - last_close)
rep["exit_after_barrier_ns"] = stats(hdr[:, 3] - hdr[:, 5])
rep["blocks"] = int(hdr.shape[0])
# hand-over between launches: how many blocks of the last launch began before the previous launch had fully exited
prev_end = int(hdr_p[:, 2].max())
rep["blocks_started_before_prev_launch_ended"] = int((hdr[:, 0] < prev_end).sum())
rep["entry_minus_prev_launch_end_ns"] = {"min": int((hdr[:, 0] - prev_end).min()), "median": float(np.median(hdr[:, 0] - prev_end)), "max": int((hdr[:, 0] - prev_end).max())}
# per SM: idle time between a block of the previous launch leaving and a block of this launch entering
gaps = []
for sm in np.unique(hdr[:, 4]):
    ex = np.sort(hdr_p[hdr_p[:, 4] == sm, 2]); en = np.sort(hdr[hdr[:, 4] == sm, 0])
    for a_, b_ in zip(ex, en):
        gaps.append(b_ - a_)
rep["per_sm_slot_handover_ns"] = {"n": len(gaps), "median": float(np.median(gaps)), "p90": float(np.percentile(gaps, 90)), "max": int(max(gaps))} if gaps else {}
per_round = {}
for jj in range(8):
    v = valid[:, :, jj]
    if v.any():
        per_round[jj] = {"landed_after_entry_ns": stats((maps[:, :, jj, 1] - hdr[:, 1][:, None])[v]),
                         "wait_ns": stats((maps[:, :, jj, 1] - maps[:, :, jj, 0])[v])}
rep["per_round"] = per_round
print(json.dumps(rep, indent=1))
if args.out:
    with open(args.out, "w") as f:
        json.dump(rep, f, indent=1)
