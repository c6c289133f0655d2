#!/usr/bin/env python
"""bench.py - heatmaps/s of the fused heatmap hot path (gen + loss + decode + PCK) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload NAME]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
           --master-port P bench.py --gpus N --steps K --warmup W

Metric (BASELINE.json): heatmaps/s of gen+loss+decode+PCK over 21x64x64 fp32 maps.  A "step" is one
pass of the hot path over one batch: configs[1] = 256 samples x 21 joints per GPU (weak scaling,
batch-sharded; one NCCL all-reduce of 46 doubles per step when N > 1).

One JSON line on rank 0:
  value      device-timed (CUDA events, max over ranks), inputs resident in HBM
  e2e        same metric through the public host-buffer API (pinned host -> H2D -> kernel -> D2H)
  roofline   the dominant kernel's algorithmic bytes / its event-timed duration vs measured HBM peak
  cpu_baseline  the oracle port of the reference path on this box's host cores (bounded sample)
`--impl reference` times that CPU path as the reference arm (the reference is pure Python: the
oracle port - validated bit-exact against the real reference - is what can travel to the GPU box).
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
PKG = "domain-adaptative-hand-pose-estimation_b200"

K_JOINTS = 21
METRIC = "heatmaps/sec (21x64x64 gen+loss+decode+PCK)"
UNIT = "heatmaps/s"
KL_EPS = 1e-7
N_SETS = 8                      # rotating input sets: 8 x 88 MB = 704 MB >> 126 MB L2
FALLBACK_HBM_GBS = 6650.0       # B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent

WORKLOADS = {
    # name: (per-GPU batch, heatmap side, description)
    "pipeline64": (256, 64, "configs[1]: 1xB200 full heatmap gen+loss(MSE+KL)+decode+PCK pipeline, "
                            "batch 256x21x64x64 fp32 per GPU"),
    "pipeline128": (1024, 128, "configs[4] shape: 21x128x128 end-to-end heatmap pipeline, batch 1024 per GPU"),
}


def algorithmic_bytes_per_map(side):
    """read pred (H*W*4) + joint f64x2, vis, weight (24) + coords out (8)."""
    return side * side * 4 + 24 + 8


# ------------------------------------------------------------------------------------------------
# clocks sampler (B200_PROFILING.md: sample DURING the timed region)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index, period_ms=50):
        self.gpu = gpu_index
        self.samples = []          # (t, sm_mhz, max_mhz, reasons)
        self.proc = None
        self.period_ms = period_ms
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", str(self.period_ms)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def _pump(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm, mx = float(parts[1]), float(parts[2])
            except ValueError:
                continue
            reasons = [n for n, v in zip(names, parts[5:9]) if v.lower().startswith("active")]
            self.samples.append((time.perf_counter(), sm, mx, reasons))

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def summary(self, windows):
        """median SM clock over the samples taken inside the [t0, t1] windows (GPU under load)."""
        inside = [s for s in self.samples if any(t0 <= s[0] <= t1 for t0, t1 in windows)]
        chosen, where = (inside, "timed regions") if inside else (self.samples, "whole run (timed region shorter than the sampling period)")
        if not chosen:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "window": "nvidia-smi unavailable"}
        reasons = sorted({r for s in chosen for r in s[3]})
        return {"sm_mhz": statistics.median(s[1] for s in chosen), "sm_max_mhz": max(s[2] for s in chosen),
                "reasons": reasons, "samples": len(chosen), "window": where}


# ------------------------------------------------------------------------------------------------
# CPU reference path (oracle port), shared by cpu_baseline and --impl reference
# ------------------------------------------------------------------------------------------------
def cpu_reference_setup():
    import torch
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    from oracle import hp_oracle as O
    return O, torch.get_num_threads()


def cpu_reference_step(O, batch, side):
    """One pass of the reference path on the host: generate_target x B + JointsMSELoss + JointsKLLoss +
    accuracy (2x get_max_preds + PCK)."""
    return O.pipeline(batch["pred"], batch["joints"], batch["vis"], kl_epsilon=KL_EPS,
                      image_size=(4 * side, 4 * side))


def run_cpu_baseline(side, budget_s=12.0, sample_B=32, min_reps=3, max_reps=400):
    synth = importlib.import_module(PKG + ".synth")
    O, threads = cpu_reference_setup()
    batch = synth.make_host_batch(1234, sample_B, K_JOINTS, side, side, image_size=4 * side)
    cpu_reference_step(O, batch, side)                       # warm-up (LUT-free path, torch thread pool)
    reps, t0 = 0, time.perf_counter()
    while reps < max_reps and (reps < min_reps or time.perf_counter() - t0 < budget_s):
        cpu_reference_step(O, batch, side)
        reps += 1
    dt = time.perf_counter() - t0
    return {"value": sample_B * K_JOINTS * reps / dt, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{reps} x ({sample_B}x{K_JOINTS}x{side}x{side}) batches of the bench recipe in {dt:.1f} s "
                      f"(oracle/hp_oracle.pipeline: torch ops on {threads} threads, numpy/Python stages single-threaded)"}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    per_gpu_B, side, desc = WORKLOADS[args.workload]
    synth = importlib.import_module(PKG + ".synth")
    O, threads = cpu_reference_setup()
    sample_B = 32
    batch = synth.make_host_batch(1234, sample_B, K_JOINTS, side, side, image_size=4 * side)
    for _ in range(max(1, min(args.warmup, 3))):
        cpu_reference_step(O, batch, side)
    budget = 150.0
    steps, t0 = 0, time.perf_counter()
    while steps < args.steps and time.perf_counter() - t0 < budget:
        cpu_reference_step(O, batch, side)
        steps += 1
    dt = time.perf_counter() - t0
    value = sample_B * K_JOINTS * steps / dt
    sample = (f"each step = one {sample_B}x{K_JOINTS}x{side}x{side} slice of the {per_gpu_B}-sample batch; "
              f"{steps} steps in {dt:.1f} s (capped at {budget:.0f} s)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": min(args.warmup, 3), "ms_per_step": 1e3 * dt / max(steps, 1), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": desc, "host": "CPU only (reference path has no GPU kernels of its own)",
                   "losses": "mse+kl", "kl_epsilon": KL_EPS, "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------
# product arm
# ------------------------------------------------------------------------------------------------
def measured_hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy read+write)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def recorded_traffic(workload):
    """dram bytes per launch of the dominant kernel from the committed ncu --set full capture, if any."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get(workload)
    except Exception:
        return None


def run_product_arm(args):
    import torch
    import torch.distributed as dist

    hp = importlib.import_module(PKG)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs CUDA: the heatmap path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"      # the version banner goes to stdout and would precede the JSON line
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    n_gpus = world
    if args.gpus != n_gpus and rank == 0:
        print(f"# note: --gpus {args.gpus} but WORLD_SIZE={world}; reporting n_gpus={n_gpus}", file=sys.stderr)

    per_gpu_B, side, desc = WORKLOADS[args.workload]
    K = K_JOINTS
    pipe = hp.HeatmapPipeline(num_keypoints=K, heatmap_size=(side, side), image_size=(4 * side, 4 * side),
                              sigma=2, kl_epsilon=KL_EPS, device=dev, collective=args.collective)
    map_bytes = side * side * 4
    n_sets = max(4, min(N_SETS, int(8e9 // (per_gpu_B * K * map_bytes)) or 1))
    sets = [hp.synth.make_device_batch(1234 + 1000 * 1 + 97 * rank + 7919 * s, per_gpu_B, K, side, side,
                                       image_size=4 * side, device=dev) for s in range(n_sets)]
    outs = [pipe.alloc_outputs(per_gpu_B, dev) for _ in range(n_sets)]
    torch.cuda.synchronize()

    sampler = ClockSampler(local)
    sampler.start()
    windows = []

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, profile=False):
        """barrier+sync, CUDA events on the launching stream around exactly `steps` calls, barrier+sync;
        -> max over ranks of the elapsed milliseconds.  `profile` brackets the region with cudaProfilerStart/Stop
        so `ncu --profile-from-start off` lists exactly the launches of the timed steps (a no-op otherwise)."""
        barrier()
        if profile:
            torch.cuda.profiler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0 = time.perf_counter()
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        if profile:
            torch.cuda.profiler.stop()
        windows.append((w0, time.perf_counter()))
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # steps are independent resident batches with their own output buffers: consecutive launches may overlap
    # (programmatic dependent launch, HP_PIPE_OVERLAP_PREV); --no-overlap serialises them
    overlap = not args.no_overlap

    def step(i):
        s = sets[i % n_sets]
        pipe(s["pred"], s["joints"], s["vis"], out=outs[i % n_sets], overlap=overlap)

    def steps_then_join(i):
        step(i)
        if i == args.steps - 1:
            pipe.join()          # the timed region ends only after the last collective + finalise

    warm = max(args.warmup, 3)
    for i in range(warm):
        step(i)
    pipe.join()
    barrier()

    # Optional: replay the n_sets-step train from a CUDA graph (--graph on).  Measured on B200: slower than eager
    # launches (15.5 vs 12.6 us per step at N=1) because a graph does not keep the programmatic overlap between
    # consecutive launches across replays; kept as an experiment switch, off by default.
    graph, graph_note = None, "eager launches"
    use_graph = args.graph == "on"
    if use_graph and args.steps >= n_sets:
        try:
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                for s_i in range(n_sets):
                    step(s_i)
            for _ in range(3):
                graph.replay()
            torch.cuda.synchronize()
            graph_note = f"CUDA graph of {n_sets} steps, replayed"
        except Exception as exc:  # noqa: BLE001 - fall back to eager launches and say so in the line
            graph = None
            graph_note = f"eager launches (graph capture failed: {type(exc).__name__})"
            torch.cuda.synchronize()
    barrier()
    if graph is not None:
        n_replays, rest = divmod(args.steps, n_sets)

        def run_steps(i):
            if i < n_replays:
                graph.replay()
            else:
                for r in range(rest):
                    step(r)
                pipe.join()
        ms_total = timed(run_steps, n_replays + 1, profile=True)
    else:
        ms_total = timed(steps_then_join, args.steps, profile=True)
    maps_per_step = per_gpu_B * K * n_gpus
    value = maps_per_step * args.steps / (ms_total * 1e-3)
    # dominant kernel alone (identical to the step at N=1; without the collective at N>1)
    def kernel_only(i):
        s = sets[i % n_sets]
        pipe.launch_local(s["pred"], s["joints"], s["vis"], outs[i % n_sets], overlap=overlap)

    for i in range(3):
        kernel_only(i)
    ms_kernel = timed(kernel_only, args.steps) / args.steps

    # the same launches fully serialised (no programmatic dependent launch): the latency of ONE launch incl. its
    # start-up, drain and the launch gap - reported beside the throughput figure, not instead of it
    def kernel_serial(i):
        s = sets[i % n_sets]
        pipe.launch_local(s["pred"], s["joints"], s["vis"], outs[i % n_sets], overlap=False)

    for i in range(3):
        kernel_serial(i)
    ms_serial = timed(kernel_serial, max(100, args.steps // 4)) / max(100, args.steps // 4)
    peak, peak_src = measured_hbm_peak()
    alg_bytes = algorithmic_bytes_per_map(side) * per_gpu_B * K
    achieved = alg_bytes / (ms_kernel * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": recorded_traffic(args.workload), "kernel": "hp::pipeline_bulk_kernel",
                "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": ms_kernel, "peak_source": peak_src,
                "kernel_ms_note": "average over the timed train of launches (CUDA events on the launching stream)"
                                  + ("; up to 4 launches of the train are resident at once" if overlap else ""),
                "serialised": {"kernel_ms": ms_serial, "achieved": alg_bytes / (ms_serial * 1e-3) / 1e9,
                               "frac": alg_bytes / (ms_serial * 1e-3) / 1e9 / peak,
                               "note": "one launch at a time (launch gap, start-up and drain exposed)"}}

    # end to end through the public host-buffer API
    host_sets = []
    for s in range(2):
        hb = hp.synth.make_host_batch(4321 + 31 * rank + s, per_gpu_B, K, side, side, image_size=4 * side)
        host_sets.append({k: torch.from_numpy(v).pin_memory() for k, v in hb.items()})
    e2e_steps = max(3, min(args.steps, 50))
    for i in range(2):
        hs = host_sets[i % 2]
        pipe.run_host(hs["pred"], hs["joints"], hs["vis"], slab=args.slab, want_pred_xy=False)
    barrier()
    w0 = time.perf_counter()
    for i in range(e2e_steps):
        hs = host_sets[i % 2]
        r = pipe.run_host(hs["pred"], hs["joints"], hs["vis"], slab=args.slab, want_pred_xy=False)
    local_dt = time.perf_counter() - w0
    barrier()
    windows.append((w0, time.perf_counter()))
    dt_t = torch.tensor([local_dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(dt_t, op=dist.ReduceOp.MAX)
    e2e_dt = float(dt_t.item())
    h2d = per_gpu_B * K * map_bytes + per_gpu_B * K * 20
    d2h = (4 + K) * 8
    e2e = {"value": maps_per_step * e2e_steps / e2e_dt, "unit": UNIT, "h2d_bytes_per_step": h2d * n_gpus,
           "d2h_bytes_per_step": d2h * n_gpus, "steps": e2e_steps, "ms_per_step": 1e3 * e2e_dt / e2e_steps,
           "api": "HeatmapPipeline.run_host -> hp_pipeline_fused_host (pinned host buffers, slabbed H2D overlapped "
                  "with the kernel, D2H of the 25-double result)", "check_avg_acc": r["avg_acc"]}
    sampler.stop()
    clocks = sampler.summary(windows)

    cpu_baseline = None
    if rank == 0 and n_gpus == 1 and not args.no_cpu_baseline:
        cpu_baseline = run_cpu_baseline(side)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps, "warmup": warm,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc, "per_gpu_batch": per_gpu_B, "joints": K, "heatmap": [side, side],
                       "losses": "mse+kl", "kl_epsilon": KL_EPS,
                       "l2": f"{n_sets} rotating input sets of {per_gpu_B * K * map_bytes / 1e6:.0f} MB "
                             f"({n_sets * per_gpu_B * K * map_bytes / 1e6:.0f} MB > 126 MB L2)",
                       "parallelism": (f"batch-sharded dp{n_gpus}; per step one exchange of {4 + 2 * K + 6} int64 "
                                       + ("over NVLink peer memory inside the fused kernel's last block (one kernel per step, "
                                          "no NCCL on the step path)"
                                          if args.collective == "peer" else "by NCCL all-reduce"))
                                      if n_gpus > 1 else "single GPU, no collective",
                       "launch": graph_note + (", consecutive steps overlap by programmatic dependent launch "
                                               "(independent resident batches, separate outputs)" if overlap else
                                               ", fully serialised")},
            "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "clocks": clocks,
            "gpu_launches": args.steps * (2 if (n_gpus > 1 and args.collective == "nccl") else 1),
        }
        sys.stdout.write(json.dumps(line) + "\n")
        sys.stdout.flush()
    if world > 1:
        pipe.close()
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20000)
    ap.add_argument("--warmup", type=int, default=200)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="pipeline64", choices=sorted(WORKLOADS))
    ap.add_argument("--slab", type=int, default=32, help="samples per H2D slab in the end-to-end path")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--graph", default="off", choices=["on", "off"],
                    help="replay the step train from a CUDA graph (measured slower: a graph serialises the launch train)")
    ap.add_argument("--collective", default="peer", choices=["peer", "nccl"],
                    help="N>1: how the ranks' partial vectors are summed each step")
    ap.add_argument("--no-overlap", action="store_true",
                    help="launch every step fully serialised after the previous one (no programmatic dependent launch)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_product_arm(args)


if __name__ == "__main__":
    sys.exit(main())
