#!/usr/bin/env python
"""bench.py - heatmaps/s of the fused heatmap hot path (gen + loss + decode + PCK) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload NAME]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
           --master-port P bench.py --gpus N --steps K --warmup W

Metric (BASELINE.json): heatmaps/s of gen+loss+decode+PCK over 21x64x64 fp32 maps.  A "step" is one
pass of the hot path over one batch.  Workloads (`--workload`; the default is the one the metric is quoted on):

  pipeline64        configs[1]  256 x 21 x 64 x 64 per GPU (weak scaling), gen + MSE + KL + decode + PCK, one kernel
  regdisp512        configs[2]  512 x 21 x 64 x 64 per GPU, pseudo-label + KL regression disparity (x6)
  fuse2048          configs[3]  2048 samples over the N GPUs (strong scaling), fuse 32/64/128 + decode + PCK
  pipeline128x8192  configs[4]  8192 x 21 x 128 x 128 over the N GPUs (strong scaling), the pipeline64 kernel at 128x128
  pipeline128       configs[4] shape at a fixed 1024 samples per GPU (weak scaling)

One JSON line on rank 0:
  value      device-timed (CUDA events, max over ranks), inputs resident in HBM: the MEDIAN of R timed regions
             of exactly K steps each (R chosen so that the regions add up to >= 250 ms: a 20-step region of a
             13 us kernel is shorter than any clock sampler's period); the R raw values are in `regions`
  e2e        same metric through the public host-buffer API (pinned host -> H2D -> kernel -> D2H)
  roofline   the dominant kernel's algorithmic bytes / its event-timed duration vs measured HBM peak
  cpu_baseline  the CPU path of the reference on this box's host cores (bounded sample) + a parity check of the
             GPU result against it on the same batch
  parity_check  N > 1: the sharded totals equal a single-GPU run of the concatenated batch bit for bit (untimed)
`--impl reference` times the reference's own CPU implementation: the REAL reference modules when a copy travels with
the repo (oracle/_ref/reference, made by __graft_entry__.build()), else the oracle port (pinned bit-equal to it).
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
MIN_TIMED_MS = 250.0            # the R regions of K steps add up to at least this much device time
MAX_REGIONS = 400

WORKLOADS = {
    "pipeline64": dict(kind="pipeline", per_gpu_B=256, side=64, scaling="weak", metric=METRIC,
                       desc="configs[1]: 1xB200 full heatmap gen+loss(MSE+KL)+decode+PCK pipeline, "
                            "batch 256x21x64x64 fp32 per GPU"),
    "pipeline128": dict(kind="pipeline", per_gpu_B=1024, side=128, scaling="weak",
                        metric="heatmaps/sec (21x128x128 gen+loss+decode+PCK)",
                        desc="configs[4] shape: 21x128x128 end-to-end heatmap pipeline, batch 1024 per GPU"),
    "pipeline128x8192": dict(kind="pipeline", total_B=8192, side=128, scaling="strong",
                             metric="heatmaps/sec (21x128x128 gen+loss+decode+PCK)",
                             desc="configs[4]: batch 8192x21x128x128 end-to-end heatmap pipeline, the FIXED batch "
                                  "sharded over the GPUs"),
    "regdisp512": dict(kind="regdisp", per_gpu_B=512, side=64, scaling="weak",
                       metric="heatmaps/sec (21x64x64 pseudo-label + KL regression disparity)",
                       desc="configs[2]: RegDA pseudo-label + JointsKLLoss regression disparity (x6), "
                            "batch 512x21x64x64 per GPU"),
    "fuse2048": dict(kind="fuse", total_B=2048, side=128, scaling="strong",
                     metric="heatmaps/sec (fuse 32/64/128 + decode + PCK)",
                     desc="configs[3]: multiscale 32/64/128 heatmap fusion + decode + PCK, batch 2048 sharded over "
                          "the GPUs, one exchange of the PCK counts"),
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

    def __init__(self, gpu_index, period_ms=20):
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
        chosen, where = (inside, "timed regions") if inside else (self.samples, "whole run (no sample fell inside a timed region)")
        if not chosen:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "window": "nvidia-smi unavailable"}
        reasons = sorted({r for s in chosen for r in s[3]})
        return {"sm_mhz": statistics.median(s[1] for s in chosen), "sm_max_mhz": max(s[2] for s in chosen),
                "reasons": reasons, "samples": len(chosen), "window": where}


# ------------------------------------------------------------------------------------------------
# CPU reference path, shared by cpu_baseline and --impl reference
# ------------------------------------------------------------------------------------------------
def shipped_reference_dir():
    """A copy of the reference's hot-path Python files that travels with the repo (git-ignored, made by
    __graft_entry__.build() from /root/reference where that exists), or the tree itself in the build container."""
    for cand in (os.environ.get("HP_REF_DIR"), os.path.join(ROOT, "oracle", "_ref", "reference"), "/root/reference"):
        if cand and os.path.isfile(os.path.join(cand, "utils", "keypoint_detection.py")):
            return cand
    return None


class CpuReference:
    """The reference path on the host.  kind == "reference": the reference's OWN functions (loaded in place by
    oracle/ref_loader.py with import shims only); kind == "port": oracle/hp_oracle.py (pinned bit-equal to them)."""

    def __init__(self, prefer_real=True):
        import torch
        self.torch = torch
        self.threads = os.cpu_count() or 1
        torch.set_num_threads(self.threads)
        self.threads = torch.get_num_threads()
        from oracle import hp_oracle as O
        self.O = O
        self.kind, self.ref = "port", None
        ref_dir = shipped_reference_dir() if prefer_real else None
        if ref_dir is not None:
            try:
                import warnings
                warnings.filterwarnings("ignore", category=SyntaxWarning)   # the reference's docstrings ('\\l' escapes)
                os.environ["HP_REF_DIR"] = ref_dir
                from oracle import ref_loader
                self.ref = ref_loader.load()
                self.kind = "reference"
            except Exception as exc:  # noqa: BLE001 - a broken copy must not take the bench down: use the port, say so
                print(f"# reference copy at {ref_dir} failed to load ({type(exc).__name__}: {exc}); using the oracle port",
                      file=sys.stderr)
        self._mods = {}

    def describe(self):
        return ("the reference's own modules, loaded in place" if self.kind == "reference"
                else "oracle/hp_oracle.py port of the reference path")

    # -- configs[0]/[1]: gen + MSE + KL + decode + PCK -----------------------------------------------------------
    def pipeline(self, batch, side):
        import numpy as np
        torch = self.torch
        image = (4 * side, 4 * side)
        if self.ref is None:
            return self.O.pipeline(batch["pred"], batch["joints"], batch["vis"], kl_epsilon=KL_EPS, image_size=image)
        R = self.ref
        pred, joints, vis = batch["pred"], batch["joints"], batch["vis"]
        B = pred.shape[0]
        tw = [R.generate_target(joints[b], vis[b], (side, side), 2, image) for b in range(B)]   # dataset side, per sample
        target = np.stack([t for t, _ in tw])
        weight = np.stack([w for _, w in tw])
        tp, tt, wt = torch.from_numpy(pred), torch.from_numpy(target), torch.from_numpy(weight)
        if "mse" not in self._mods:
            self._mods["mse"], self._mods["kl"] = R.JointsMSELoss(), R.JointsKLLoss(epsilon=KL_EPS)
        mse = self._mods["mse"](tp, tt, wt)
        kl = self._mods["kl"](tp, tt, wt)
        acc, avg, cnt, xy = R.accuracy(pred, target)
        return dict(mse=float(mse), kl=float(kl), acc=acc, avg_acc=float(avg), cnt=int(cnt), pred_xy=xy)

    # -- configs[2]: pseudo-label + KL regression disparity (x6), 'min' then 'max' ------------------------------
    def regdisp(self, y, y_adv, mode):
        torch = self.torch
        ty, ta = torch.from_numpy(y), torch.from_numpy(y_adv)
        if self.ref is None:
            return float(self.O.regression_disparity("x6", ty, ta, None, None, mode, KL_EPS))
        R = self.ref
        if "rd6" not in self._mods:
            self._mods["rd6"] = R.RegressionDisparityx6(R.PseudoLabelGenerator(K_JOINTS, y.shape[2], y.shape[3]),
                                                        R.JointsKLLoss(epsilon=KL_EPS))
        return float(self._mods["rd6"](ty, ta, None, None, mode))

    # -- configs[3]: fuse 32/64/128 + decode + PCK ----------------------------------------------------------------
    def fuse(self, lo, mid, hi, target_xy):
        import numpy as np
        torch = self.torch
        up = torch.nn.functional.interpolate
        H, W = hi.shape[2], hi.shape[3]
        fused = (0.5 * up(torch.from_numpy(lo), size=(H, W), mode="bilinear") +
                 up(torch.from_numpy(mid), size=(H, W), mode="bilinear") + torch.from_numpy(hi)).numpy()
        gmp = self.ref.get_max_preds if self.ref is not None else self.O.get_max_preds
        xy, _ = gmp(fused)
        hits, valid = self.O.pck_counts(xy, target_xy.astype(np.float32), H, W, 0.5)
        return xy, hits, valid


def host_batch_for(wl, seed, B):
    import numpy as np
    synth = importlib.import_module(PKG + ".synth")
    side = wl["side"]
    d = synth.make_host_batch(seed, B, K_JOINTS, side, side, image_size=4 * side)
    if wl["kind"] == "regdisp":
        d["y_adv"] = synth.make_host_batch(seed + 5000, B, K_JOINTS, side, side, image_size=4 * side)["pred"]
    elif wl["kind"] == "fuse":
        hi = d["pred"]
        d["mid"] = hi.reshape(B, K_JOINTS, side // 2, 2, side // 2, 2).mean(axis=(3, 5)).astype(np.float32)
        d["lo"] = hi.reshape(B, K_JOINTS, side // 4, 4, side // 4, 4).mean(axis=(3, 5)).astype(np.float32)
        d["target_xy"] = np.random.RandomState(seed + 77).randint(0, side, size=(B, K_JOINTS, 2)).astype(np.float32)
    return d


def cpu_step(cpu, wl, batch):
    if wl["kind"] == "pipeline":
        return cpu.pipeline(batch, wl["side"])
    if wl["kind"] == "regdisp":
        return cpu.regdisp(batch["pred"], batch["y_adv"], "min")
    return cpu.fuse(batch["lo"], batch["mid"], batch["pred"], batch["target_xy"])


def run_cpu_baseline(wl, gpu_check=None, budget_s=12.0, sample_B=32, min_reps=3, max_reps=400):
    cpu = CpuReference()
    side = wl["side"]
    batch = host_batch_for(wl, 1234, sample_B)
    want = cpu_step(cpu, wl, batch)                               # warm-up (torch thread pool, LUT of the real PLG)
    reps, t0 = 0, time.perf_counter()
    while reps < max_reps and (reps < min_reps or time.perf_counter() - t0 < budget_s):
        cpu_step(cpu, wl, batch)
        reps += 1
    dt = time.perf_counter() - t0
    out = {"value": sample_B * K_JOINTS * reps / dt, "unit": UNIT, "cores": cpu.threads, "kind": cpu.kind,
           "sample": f"{reps} x ({sample_B}x{K_JOINTS}x{side}x{side}) batches of the bench recipe in {dt:.1f} s "
                     f"({cpu.describe()}: torch ops on {cpu.threads} threads, numpy/Python stages single-threaded)"}
    if gpu_check is not None:
        try:
            out["parity"] = gpu_check(batch, want)
        except Exception as exc:  # noqa: BLE001 - report, never hide
            out["parity"] = {"ok": False, "error": f"{type(exc).__name__}: {exc}"}
    return out


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    wl = WORKLOADS[args.workload]
    side = wl["side"]
    cpu = CpuReference()
    sample_B = 32
    batch = host_batch_for(wl, 1234, sample_B)
    warm = max(1, min(args.warmup, 3))
    for _ in range(warm):
        cpu_step(cpu, wl, batch)
    budget = 150.0
    steps, t0 = 0, time.perf_counter()
    while steps < args.steps and time.perf_counter() - t0 < budget:
        cpu_step(cpu, wl, batch)
        steps += 1
    dt = time.perf_counter() - t0
    value = sample_B * K_JOINTS * steps / dt
    sample = (f"each step = one {sample_B}x{K_JOINTS}x{side}x{side} slice of the workload's batch; "
              f"{steps} steps in {dt:.1f} s (capped at {budget:.0f} s); {cpu.describe()}")
    line = {
        "impl": "reference", "metric": wl["metric"], "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": 1e3 * dt / max(steps, 1), "higher_is_better": True,
        "scaling": wl["scaling"], "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["desc"], "host": "CPU only (reference path has no GPU kernels of its own)",
                   "losses": "mse+kl" if wl["kind"] == "pipeline" else None, "kl_epsilon": KL_EPS, "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cpu.threads, "kind": cpu.kind, "sample": sample},
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


class Harness:
    """Process-group plumbing + the timing protocol shared by the workloads."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.args = torch, dist, args
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise RuntimeError("bench.py needs CUDA: the heatmap path has no CPU fallback (use --impl reference for the CPU arm)")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        # one process per GPU: run (and first-touch the pinned staging buffers) on the GPU's own NUMA node
        self._affinity0 = os.sched_getaffinity(0) if hasattr(os, "sched_getaffinity") else None
        self.numa_node = None
        if not args.no_numa_bind:
            self.numa_node = importlib.import_module(PKG + ".dist").bind_to_gpu_numa(self.local)
        if self.world > 1:
            if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
                os.environ["NCCL_DEBUG"] = "WARN"      # the version banner goes to stdout and would precede the JSON line
            dist.init_process_group("nccl", rank=self.rank, world_size=self.world, device_id=self.dev)
        if args.gpus != self.world and self.rank == 0:
            print(f"# note: --gpus {args.gpus} but WORLD_SIZE={self.world}; reporting n_gpus={self.world}", file=sys.stderr)
        self.sampler = ClockSampler(self.local)
        self.sampler.start()
        self.windows = []

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def timed(self, fn, steps, profile=False):
        """barrier+sync, CUDA events on the launching stream around exactly `steps` calls, barrier+sync;
        -> max over ranks of the elapsed milliseconds.  `profile` brackets the region with cudaProfilerStart/Stop
        so `ncu --profile-from-start off` lists exactly the launches of the timed steps (a no-op otherwise)."""
        torch = self.torch
        self.barrier()
        if profile:
            torch.cuda.profiler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0 = time.perf_counter()
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        self.barrier()
        if profile:
            torch.cuda.profiler.stop()
        self.windows.append((w0, time.perf_counter()))
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(ms, op=self.dist.ReduceOp.MAX)
        return float(ms.item())

    def regions(self, fn, steps, min_total_ms=MIN_TIMED_MS, profile_first=False):
        """R timed regions of exactly `steps` steps each (every one bracketed like `timed`), until they add up to
        `min_total_ms` of device time (>= 3, <= MAX_REGIONS).  The values are max-over-ranks, so every rank takes the
        same decision.  -> list of ms per region.  The clock sampler sees ONE window over all R regions: the GPU is
        under the timed load for most of it, and it is long enough to hold several nvidia-smi samples."""
        first = len(self.windows)
        out = []
        while True:
            out.append(self.timed(fn, steps, profile=profile_first and not out))
            if len(out) >= MAX_REGIONS or (len(out) >= 3 and sum(out) >= min_total_ms):
                break
        span = (self.windows[first][0], self.windows[-1][1])
        del self.windows[first:]
        self.windows.append(span)
        return out

    def finish(self):
        self.sampler.stop()
        if self._affinity0 is not None:          # the CPU baseline that may follow uses every host core again
            try:
                os.sched_setaffinity(0, self._affinity0)
            except OSError:
                pass
        return self.sampler.summary(self.windows)

    def shutdown(self):
        if self.world > 1:
            self.dist.barrier()
            self.dist.destroy_process_group()


def region_stats(ms_list, steps):
    per = sorted(m / steps for m in ms_list)
    return {"repeats": len(per), "ms_per_step_median": statistics.median(per), "ms_per_step_min": per[0],
            "ms_per_step_max": per[-1], "ms_per_step_all": [round(m / steps, 6) for m in ms_list][:64]}


def shard_B(wl, world, rank):
    if "total_B" in wl:
        base, rem = divmod(wl["total_B"], world)
        return base + (1 if rank < rem else 0)
    return wl["per_gpu_B"]


def total_B(wl, world):
    return wl["total_B"] if "total_B" in wl else wl["per_gpu_B"] * world


# ---- kind == "pipeline" ----------------------------------------------------------------------------------------------
def run_pipeline_arm(args, wl):
    h = Harness(args)
    torch, dist, dev, world, rank = h.torch, h.dist, h.dev, h.world, h.rank
    hp = importlib.import_module(PKG)
    side, K = wl["side"], K_JOINTS
    B = shard_B(wl, world, rank)
    pipe = hp.HeatmapPipeline(num_keypoints=K, heatmap_size=(side, side), image_size=(4 * side, 4 * side),
                              sigma=2, kl_epsilon=KL_EPS, device=dev, collective=args.collective)
    map_bytes = side * side * 4
    set_bytes = B * K * map_bytes
    # rotating input sets so that a step never finds its input in the 126 MB L2 (a single set larger than 4x L2 needs none)
    n_sets = 1 if set_bytes > 512e6 else max(4, min(N_SETS, int(8e9 // set_bytes) or 1))
    sets = [hp.synth.make_device_batch(1234 + 1000 * 1 + 97 * rank + 7919 * s, B, K, side, side,
                                       image_size=4 * side, device=dev) for s in range(n_sets)]
    n_outs = max(n_sets, 8)       # consecutive launches of a train need different output buffers
    outs = [pipe.alloc_outputs(B, dev) for _ in range(n_outs)]
    torch.cuda.synchronize()
    overlap = not args.no_overlap
    sharded = world > 1

    # pre-bound launches (the public plan API): what a caller that streams batches through resident buffers uses
    def make_launch(i, ov, local_only=False):
        s = sets[i % n_sets]
        if sharded and not local_only and args.collective == "peer":
            return pipe.plan_peer(s["pred"], s["joints"], s["vis"], out=outs[i % n_outs], overlap=ov)[0]
        return pipe.plan(s["pred"], s["joints"], s["vis"], out=outs[i % n_outs], finalize=True, overlap=ov)[0]

    if sharded and args.collective == "nccl":
        def step(i):
            s = sets[i % n_sets]
            pipe(s["pred"], s["joints"], s["vis"], out=outs[i % n_outs], overlap=overlap)
    else:
        launches = [make_launch(i, overlap) for i in range(n_outs)]

        def step(i):
            launches[i % n_outs]()

    def steps_then_join(i):
        step(i)
        if i == args.steps - 1:
            pipe.join()          # the timed region ends only after the last collective + finalise

    warm = max(args.warmup, 3)
    for i in range(max(warm, n_outs)):       # every pre-bound launch runs at least once before anything is timed
        step(i)
    pipe.join()
    h.barrier()

    ms_regions = h.regions(steps_then_join, args.steps, profile_first=True)
    ms_total = statistics.median(ms_regions)
    n_gpus = world
    maps_per_step = total_B(wl, world) * K
    value = maps_per_step * args.steps / (ms_total * 1e-3)

    # dominant kernel alone (identical to the step at N=1; without the collective at N>1)
    k_launch = [make_launch(i, overlap, local_only=True) for i in range(n_outs)] if sharded else None

    def kernel_only(i):
        (k_launch or launches)[i % n_outs]()

    for i in range(n_outs):
        kernel_only(i)
    ms_kernel_regions = h.regions(kernel_only, args.steps, min_total_ms=MIN_TIMED_MS / 2)
    ms_kernel = statistics.median(ms_kernel_regions) / args.steps

    # the same launches fully serialised: no data of a launch is read and nothing is written before the previous
    # launch has completed - the latency of ONE launch incl. its start-up and drain, reported beside the train figure
    s_launch = [make_launch(i, False, local_only=True) for i in range(n_outs)]

    def kernel_serial(i):
        s_launch[i % n_outs]()

    for i in range(n_outs):
        kernel_serial(i)
    n_serial = max(20, min(args.steps, 200))
    ms_serial_regions = h.regions(kernel_serial, n_serial, min_total_ms=MIN_TIMED_MS / 2)
    ms_serial = statistics.median(ms_serial_regions) / n_serial

    # one launch on an idle GPU, events right around it (includes the launch latency an isolated caller pays)
    iso = []
    for i in range(20):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        kernel_serial(i)
        e1.record()
        torch.cuda.synchronize()
        iso.append(e0.elapsed_time(e1))
    ms_isolated = statistics.median(iso)

    peak, peak_src = measured_hbm_peak()
    alg_bytes = algorithmic_bytes_per_map(side) * B * K

    def frac(ms):
        return alg_bytes / (ms * 1e-3) / 1e9 / peak

    achieved = alg_bytes / (ms_kernel * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": recorded_traffic(args.workload), "kernel": "hp::pipeline_bulk_kernel",
                "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": ms_kernel, "peak_source": peak_src,
                "kernel_ms_note": "median over the timed regions of the average launch duration in a train of launches "
                                  "(CUDA events on the launching stream)"
                                  + ("; up to 4 launches of the train are resident at once" if overlap else ""),
                "regions": region_stats(ms_kernel_regions, args.steps),
                "serialised": {"kernel_ms": ms_serial, "achieved": alg_bytes / (ms_serial * 1e-3) / 1e9, "frac": frac(ms_serial),
                               "regions": region_stats(ms_serial_regions, n_serial),
                               "note": "back-to-back launches, each ordered after the COMPLETION of the previous one "
                                       "(start-up, drain and the launch gap exposed)"},
                "isolated": {"kernel_ms": ms_isolated, "frac": frac(ms_isolated),
                             "note": "one launch on an idle GPU between two events (median of 20)"}}

    # sharded parity: all-rank totals == one GPU on the concatenated batch, bit for bit (untimed)
    parity = sharded_parity_check(h, hp, pipe, sets[0], side) if sharded else None

    # end to end through the public host-buffer API
    host_sets = []
    e2e_B = B if set_bytes <= 2e9 else max(1, int(2e9 // (K * map_bytes)))       # bound the pinned staging memory
    for s in range(2):
        hb = hp.synth.make_host_batch(4321 + 31 * rank + s, e2e_B, K, side, side, image_size=4 * side)
        host_sets.append({k: torch.from_numpy(v).pin_memory() for k, v in hb.items()})
    e2e_steps = max(3, min(args.steps, 50 if set_bytes < 512e6 else 5))
    for i in range(2):
        hs = host_sets[i % 2]
        pipe.run_host(hs["pred"], hs["joints"], hs["vis"], slab=args.slab, want_pred_xy=True)
    h.barrier()
    w0 = time.perf_counter()
    for i in range(e2e_steps):
        hs = host_sets[i % 2]
        r = pipe.run_host(hs["pred"], hs["joints"], hs["vis"], slab=args.slab, want_pred_xy=True)
    local_dt = time.perf_counter() - w0
    h.barrier()
    h.windows.append((w0, time.perf_counter()))
    dt_t = torch.tensor([local_dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(dt_t, op=dist.ReduceOp.MAX)
    e2e_dt = float(dt_t.item())
    h2d, d2h = pipe.host_bytes_per_call(e2e_B)
    e2e_maps = e2e_B * K * n_gpus
    e2e = {"value": e2e_maps * e2e_steps / e2e_dt, "unit": UNIT, "h2d_bytes_per_step": h2d * n_gpus,
           "d2h_bytes_per_step": d2h * n_gpus, "steps": e2e_steps, "ms_per_step": 1e3 * e2e_dt / e2e_steps,
           "per_gpu_batch": e2e_B, "host_numa_node": h.numa_node,
           "api": "HeatmapPipeline.run_host -> hp_pipeline_fused_host (pinned host buffers, slabbed H2D overlapped "
                  "with the kernel, D2H of the 25-double result and the decoded coordinates"
                  + ("; the ranks' partial vectors are exchanged, every rank returns the all-rank totals)" if sharded else ")"),
           "check_avg_acc": r["avg_acc"]}
    clocks = h.finish()

    cpu_baseline = None
    if rank == 0 and n_gpus == 1 and not args.no_cpu_baseline:
        def gpu_check(batch, want):
            import numpy as np
            got = pipe.run_host(batch["pred"], batch["joints"], batch["vis"], slab=args.slab, want_pred_xy=True)
            ok_xy = bool(np.array_equal(got["pred_xy"], want["pred_xy"]))
            ok_acc = bool(np.array_equal(got["acc"], want["acc"]) and got["cnt"] == want["cnt"])
            rel = {k: abs(got[k] - want[k]) / max(abs(want[k]), 1e-30) for k in ("mse", "kl")}
            return {"ok": ok_xy and ok_acc and all(v <= 1e-5 for v in rel.values()), "pred_xy_bit_equal": ok_xy,
                    "pck_bit_equal": ok_acc, "mse_rel_err": rel["mse"], "kl_rel_err": rel["kl"],
                    "what": "GPU pipeline (run_host) vs the CPU baseline's own result on its 32-sample batch"}
        cpu_baseline = run_cpu_baseline(wl, gpu_check)

    if rank == 0:
        line = {
            "metric": wl["metric"], "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps, "warmup": warm,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": wl["scaling"], "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["desc"], "per_gpu_batch": B, "joints": K, "heatmap": [side, side],
                       "losses": "mse+kl", "kl_epsilon": KL_EPS,
                       "timing": f"value = median of {len(ms_regions)} timed regions of exactly {args.steps} steps each "
                                 f"(barrier + synchronize + CUDA events around every region, max over ranks)",
                       "l2": (f"{n_sets} rotating input sets of {set_bytes / 1e6:.0f} MB "
                              f"({n_sets * set_bytes / 1e6:.0f} MB > 126 MB L2)") if n_sets > 1 else
                             f"one input set of {set_bytes / 1e6:.0f} MB per GPU (> 4x the 126 MB L2)",
                       "parallelism": (f"batch-sharded dp{n_gpus}; per step one exchange of {4 + 2 * K + 6} int64 "
                                       + ("over NVLink peer memory inside the fused kernel (one kernel per step, "
                                          "no NCCL on the step path)"
                                          if args.collective == "peer" else "by NCCL all-reduce"))
                                      if n_gpus > 1 else "single GPU, no collective",
                       "launch": "pre-bound launches (HeatmapPipeline.plan)"
                                 + (", consecutive steps overlap by programmatic dependent launch "
                                    "(independent resident batches, separate outputs)" if overlap else ", fully serialised")},
            "regions": region_stats(ms_regions, args.steps),
            "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "clocks": clocks,
            "parity_check": parity,
            "gpu_launches": args.steps * (2 if (n_gpus > 1 and args.collective == "nccl") else 1),
        }
        sys.stdout.write(json.dumps(line) + "\n")
        sys.stdout.flush()
    if world > 1:
        pipe.close()
    h.shutdown()
    return 0


def sharded_parity_check(h, hp, pipe, s, side):
    """Untimed: run the SHARDED step on (a slice of) every rank's first input set, gather the slices on every rank,
    run rank 0's single-GPU pipeline on the concatenation and compare bit for bit."""
    torch, dist, dev, world, rank = h.torch, h.dist, h.dev, h.world, h.rank
    K = K_JOINTS
    cB = min(s["pred"].shape[0], 256 if side <= 64 else 64)
    pred, joints, vis = (s[k][:cB].contiguous() for k in ("pred", "joints", "vis"))
    out = pipe(pred, joints, vis)
    pipe.join()
    out.wait()
    torch.cuda.synchronize()
    gathered = {}
    for name, t in (("pred", pred), ("joints", joints), ("vis", vis)):
        parts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(parts, t)
        gathered[name] = torch.cat(parts, 0)
    ok = True
    detail = {}
    if rank == 0:
        single = hp.HeatmapPipeline(num_keypoints=K, heatmap_size=(side, side), image_size=(4 * side, 4 * side),
                                    sigma=2, kl_epsilon=KL_EPS, device=dev)
        ref = single.launch_local(gathered["pred"], gathered["joints"], gathered["vis"])
        torch.cuda.synchronize()
        detail["partial_bit_equal"] = bool(torch.equal(out.partial, ref.partial))
        detail["result_bit_equal"] = bool(torch.equal(out.result.view(torch.int64), ref.result.view(torch.int64)))
        detail["pred_xy_bit_equal"] = bool(torch.equal(out.pred_xy, ref.pred_xy[:cB]))
        ok = all(detail.values())
        detail["avg_acc"] = float(ref.result[2].item())
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    # every rank must hold the same totals as rank 0
    tot = out.partial.clone()
    dist.broadcast(tot, 0)
    same = torch.tensor([1 if torch.equal(tot, out.partial) else 0], device=dev)
    dist.all_reduce(same, op=dist.ReduceOp.MIN)
    detail.update({"ok": bool(flag.item()) and bool(same.item()), "all_ranks_hold_rank0_totals": bool(same.item()),
                   "ranks": world, "samples": world * cB,
                   "what": f"sharded step over {world} ranks x {cB} samples vs rank 0's single-GPU kernel on the "
                           f"concatenated {world * cB}-sample batch: int64 partial vector, float64 result bits, decoded "
                           f"coordinates of rank 0's slice"})
    return detail


# ---- kind == "regdisp": configs[2] ------------------------------------------------------------------------------------
def run_regdisp_arm(args, wl):
    h = Harness(args)
    torch, dist, dev, world, rank = h.torch, h.dist, h.dev, h.world, h.rank
    hp = importlib.import_module(PKG)
    side, K = wl["side"], K_JOINTS
    B = shard_B(wl, world, rank)
    n, hw4 = B * K, side * side * 4
    n_sets = 3
    with torch.no_grad():
        ys = [hp.synth.make_device_batch(2000 + 97 * rank + i, B, K, side, side, image_size=4 * side, device=dev)["pred"]
              for i in range(n_sets)]
        advs = [hp.synth.make_device_batch(3000 + 97 * rank + i, B, K, side, side, image_size=4 * side, device=dev)["pred"]
                for i in range(n_sets)]
        f32 = [torch.nn.functional.avg_pool2d(a, 2) for a in advs]
        f16 = [torch.nn.functional.avg_pool2d(a, 4) for a in advs]
        t5 = [hp.fuse_multiscale(f16[i], f32[i], side, side // 2)[0] for i in range(n_sets)]
    rd6 = hp.RegressionDisparityx6(hp.PseudoLabelGenerator(K, side, side), hp.JointsKLLoss(epsilon=KL_EPS))
    variants = {
        "min": (lambda i: rd6(ys[i % n_sets], advs[i % n_sets], None, None, "min"), 2 * hw4,
                "read y (decode) + y_adv; the pseudo label is never a map (regda_7.py:3609-3632, mode='min')"),
        "max": (lambda i: rd6(ys[i % n_sets], advs[i % n_sets], None, None, "max"), 2 * hw4,
                "mode='max', y_adv2=None: ground-false label rebuilt per pixel from the sample's K centres"),
        "max_fused": (lambda i: rd6(ys[i % n_sets], advs[i % n_sets], t5[i % n_sets], None, "max"), 3 * hw4,
                      "mode='max' with the pre-fused target5 map (train1.py:419-421)"),
        "max_heads": (lambda i: rd6(ys[i % n_sets], advs[i % n_sets], heads[i % n_sets], None, "max"), 2 * hw4 + hw4 // 4 + hw4 // 16,
                      "mode='max' with target5 left unfused (hp.FusedHeads: the loss kernel interpolates 0.5 up64(y_adv3) + "
                      "up64(y_adv2) from the staged 16x16 / 32x32 heads; SURVEY.md 8d: 37,888 B per map, no fusion launch)"),
    }
    heads = [hp.FusedHeads(f16[i], f32[i]) for i in range(n_sets)]
    peak, peak_src = measured_hbm_peak()
    rows = {}
    with torch.no_grad():
        for name, (fn, bytes_per_map, note) in variants.items():
            for i in range(max(args.warmup, 3)):
                fn(i)
            reg = h.regions(fn, args.steps, min_total_ms=MIN_TIMED_MS / 2, profile_first=True)
            ms = statistics.median(reg) / args.steps
            gbs = n * bytes_per_map / (ms * 1e-3) / 1e9
            rows[name] = {"ms_per_step": ms, "heatmaps_per_s": n * world / (ms * 1e-3), "algorithmic_bytes_per_map": bytes_per_map,
                          "achieved_GBps": gbs, "frac": gbs / peak, "regions": region_stats(reg, args.steps), "note": note,
                          "launches_per_step": 1 if name == "min" else 2,
                          "traffic": recorded_traffic(args.workload + ("" if name == "min" else "_" + name))}
        # e2e: host heatmaps in, scalar loss out
        hy = [y.cpu().pin_memory() for y in ys[:2]]
        ha = [a.cpu().pin_memory() for a in advs[:2]]
        e2e_steps = max(3, min(args.steps, 20))

        def e2e_step(i):
            y = hy[i % 2].to(dev, non_blocking=True)
            a = ha[i % 2].to(dev, non_blocking=True)
            return float(rd6(y, a, None, None, "min").item())

        for i in range(2):
            e2e_step(i)
        h.barrier()
        w0 = time.perf_counter()
        for i in range(e2e_steps):
            last = e2e_step(i)
        dt = time.perf_counter() - w0
        h.barrier()
        h.windows.append((w0, time.perf_counter()))
    dt_t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(dt_t, op=dist.ReduceOp.MAX)
    clocks = h.finish()
    cpu_baseline = run_cpu_baseline(wl) if (rank == 0 and world == 1 and not args.no_cpu_baseline) else None
    if rank == 0:
        head = rows["min"]
        line = {"metric": wl["metric"], "value": head["heatmaps_per_s"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": head["ms_per_step"], "higher_is_better": True,
                "scaling": wl["scaling"], "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": wl["desc"], "per_gpu_batch": B, "joints": K, "heatmap": [side, side],
                           "step": "RegressionDisparityx6(PseudoLabelGenerator, JointsKLLoss(1e-7)) forward, mode='min' "
                                   "(decode + loss in ONE launch); 'max' variants (decode launch + dense loss launch) in `variants`",
                           "l2": f"{n_sets} rotating input sets of {2 * n * hw4 / 1e6:.0f} MB",
                           "parallelism": f"batch-sharded dp{world}, no collective (per-rank mean)" if world > 1 else "single GPU"},
                "roofline": {"bound": "hbm", "achieved": head["achieved_GBps"], "peak": peak, "unit": "GB/s", "frac": head["frac"],
                             "traffic": recorded_traffic(args.workload), "kernel": "hp::regdisp_min_kernel ('min': decode + loss in one launch); 'max' variants: hp::decode_kernel + hp::regdisp_dense_kernel",
                             "algorithmic_bytes_per_launch": n * 2 * hw4, "kernel_ms": head["ms_per_step"], "peak_source": peak_src},
                "variants": rows, "cpu_baseline": cpu_baseline,
                "e2e": {"value": n * world * e2e_steps / float(dt_t.item()), "unit": UNIT, "h2d_bytes_per_step": 2 * n * hw4 * world,
                        "d2h_bytes_per_step": 4 * world, "steps": e2e_steps, "ms_per_step": 1e3 * float(dt_t.item()) / e2e_steps,
                        "api": "pinned host y, y_adv -> .to(device) -> RegressionDisparityx6(...,'min') -> loss.item()",
                        "check_loss": last},
                "clocks": clocks, "gpu_launches": args.steps}
        sys.stdout.write(json.dumps(line) + "\n")
        sys.stdout.flush()
    h.shutdown()
    return 0


# ---- kind == "fuse": configs[3] ---------------------------------------------------------------------------------------
def run_fuse_arm(args, wl):
    h = Harness(args)
    torch, dist, dev, world, rank = h.torch, h.dist, h.dev, h.world, h.rank
    hp = importlib.import_module(PKG)
    side, K = wl["side"], K_JOINTS
    B = shard_B(wl, world, rank)
    n = B * K
    bytes_map = side * side * 4 + side * side + side * side // 4 + 16
    n_sets = 3 if n * bytes_map < 2e9 else 1
    with torch.no_grad():
        hi = [hp.synth.make_device_batch(4000 + 97 * rank + i, B, K, side, side, image_size=4 * side, device=dev)["pred"]
              for i in range(n_sets)]
        mid = [torch.nn.functional.avg_pool2d(x, 2) for x in hi]
        lo = [torch.nn.functional.avg_pool2d(x, 4) for x in hi]
        tgt = [torch.randint(0, side, (B, K, 2), device=dev).float() for _ in range(n_sets)]
        ev = hp.MultiscaleEval(K)
        fuse_mode = os.environ.get("HP_BENCH_FUSE_MODE", "sync")   # sync | deferred | twolaunch: within 1.5 % of each other at 8 GPUs

        def step(i):
            # HP_BENCH_FUSE_MODE=deferred: a train of steps whose exchange is deferred by one step (MultiscaleEval.step); the
            # region then ends only after the last step's totals have been collected (flush)
            if fuse_mode == "sync":          # the exchange completed inside every step's kernel (its last block)
                return ev(lo[i % n_sets], mid[i % n_sets], hi[i % n_sets], tgt[i % n_sets])
            if fuse_mode == "twolaunch":     # comparison runs: local kernel + the one-warp exchange kernel (first half of round 2)
                acc, xy, counts = ev(lo[i % n_sets], mid[i % n_sets], hi[i % n_sets], tgt[i % n_sets], local=True)
                if world > 1:
                    hp.dist.shared_peer_exchange(dev, None).pck_finalize(counts, K, counts, acc)
                return acc, xy, counts
            r = ev.step(lo[i % n_sets], mid[i % n_sets], hi[i % n_sets], tgt[i % n_sets])
            if i == args.steps - 1:
                ev.flush()
            return r

        for i in range(max(args.warmup, 3)):
            step(i)
        ev.flush()
        reg = h.regions(step, args.steps, profile_first=True)
        ms = statistics.median(reg) / args.steps
        # e2e
        hs = [(lo[0].cpu().pin_memory(), mid[0].cpu().pin_memory(), hi[0].cpu().pin_memory(), tgt[0].cpu().pin_memory())]
        e2e_steps = max(3, min(args.steps, 10))

        def e2e_step(i):
            a, b, c, t = (x.to(dev, non_blocking=True) for x in hs[0])
            acc, xy, _ = ev(a, b, c, t)
            return acc.cpu(), xy.cpu()

        e2e_step(0)
        h.barrier()
        w0 = time.perf_counter()
        for i in range(e2e_steps):
            acc_h, _ = e2e_step(i)
        dt = time.perf_counter() - w0
        h.barrier()
        h.windows.append((w0, time.perf_counter()))
    dt_t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(dt_t, op=dist.ReduceOp.MAX)
    parity = fuse_parity_check(h, hp, ev, lo[0], mid[0], hi[0], tgt[0]) if world > 1 else None
    clocks = h.finish()
    peak, peak_src = measured_hbm_peak()
    gbs = n * bytes_map / (ms * 1e-3) / 1e9
    cpu_baseline = run_cpu_baseline(wl) if (rank == 0 and world == 1 and not args.no_cpu_baseline) else None
    if rank == 0:
        maps = total_B(wl, world) * K
        line = {"metric": wl["metric"], "value": maps / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": wl["scaling"],
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": wl["desc"], "per_gpu_batch": B, "joints": K, "heatmap": [side, side],
                           "l2": f"{n_sets} input set(s) of {n * bytes_map / 1e6:.0f} MB per GPU",
                           "parallelism": f"batch-sharded dp{world}; one exchange of the 2K integer PCK counts per step, inside the "
                                          f"kernel's last block over NVLink peer memory ({fuse_mode})"
                                          if world > 1 else "single GPU, no collective"},
                "regions": region_stats(reg, args.steps),
                "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
                             "traffic": recorded_traffic(args.workload), "kernel": "hp::fuse_block_hs_kernel (fuse + decode + PCK; the three sources staged by the copy engine)",
                             "algorithmic_bytes_per_launch": n * bytes_map, "kernel_ms": ms, "peak_source": peak_src},
                "cpu_baseline": cpu_baseline,
                "e2e": {"value": maps * e2e_steps / float(dt_t.item()), "unit": UNIT,
                        "h2d_bytes_per_step": (n * bytes_map - 8 * n) * world, "d2h_bytes_per_step": ((K + 2) * 8 + 8 * n) * world,
                        "steps": e2e_steps, "ms_per_step": 1e3 * float(dt_t.item()) / e2e_steps,
                        "api": "pinned host lo/mid/hi/target_xy -> .to(device) -> MultiscaleEval -> acc, pred_xy .cpu()",
                        "check_avg_acc": float(acc_h[K].item())},
                "parity_check": parity,
                "clocks": clocks, "gpu_launches": args.steps}
        sys.stdout.write(json.dumps(line) + "\n")
        sys.stdout.flush()
    h.shutdown()
    return 0


def fuse_parity_check(h, hp, ev, lo, mid, hi, tgt):
    """Untimed, N > 1: the sharded fuse + decode + PCK step (counts summed over the ranks) on a slice of every rank's
    inputs against rank 0's single-GPU step on the gathered slices: integer hit / valid counts, the float64 accuracy
    vector and the decoded coordinates of rank 0's slice, bit for bit (keypoint_detection.py:63-92 on the whole batch)."""
    torch, dist, dev, world, rank = h.torch, h.dist, h.dev, h.world, h.rank
    cB = min(lo.shape[0], 64)
    parts = [t[:cB].contiguous() for t in (lo, mid, hi, tgt)]
    with torch.no_grad():
        acc, xy, counts = ev(*parts)
        torch.cuda.synchronize()
        gathered = []
        for t in parts:
            g = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(g, t)
            gathered.append(torch.cat(g, 0))
        ok, detail = True, {}
        if rank == 0:
            acc1, xy1, counts1 = ev(*gathered, local=True)
            torch.cuda.synchronize()
            detail["counts_bit_equal"] = bool(torch.equal(counts, counts1))
            detail["acc_bit_equal"] = bool(torch.equal(acc.view(torch.int64), acc1.view(torch.int64)))
            detail["pred_xy_bit_equal"] = bool(torch.equal(xy, xy1[:cB]))
            ok = all(detail.values())
            detail["avg_acc"] = float(acc1[ev.K].item())
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    tot = counts.clone()
    dist.broadcast(tot, 0)
    same = torch.tensor([1 if torch.equal(tot, counts) else 0], device=dev)
    dist.all_reduce(same, op=dist.ReduceOp.MIN)
    detail.update({"ok": bool(flag.item()) and bool(same.item()), "all_ranks_hold_rank0_totals": bool(same.item()),
                   "ranks": world, "samples": world * cB,
                   "what": f"sharded fuse + decode + PCK over {world} ranks x {cB} samples vs rank 0's single-GPU step on the "
                           f"gathered {world * cB}-sample batch: int32 hit / valid counts, float64 accuracy vector bits, "
                           f"decoded coordinates of rank 0's slice"})
    return detail


def _keep_stdout_for_the_json_line():
    """The contract is ONE JSON line on stdout.  Native libraries write there too (NCCL's "NCCL version ..." banner at the first
    communicator, whatever NCCL_DEBUG says once torch has loaded it): file descriptor 1 is pointed at stderr for everything
    below Python, and sys.stdout keeps the real stdout for the line itself."""
    try:
        sys.stdout.flush()
        real = os.dup(1)
        os.dup2(2, 1)
        sys.stdout = os.fdopen(real, "w", buffering=1)
    except OSError:
        pass


def main():
    _keep_stdout_for_the_json_line()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None, help="steps per timed region (default: 2000 for the 64x64 pipeline, "
                                                            "fewer for the larger workloads)")
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="pipeline64", choices=sorted(WORKLOADS))
    ap.add_argument("--slab", type=int, default=64, help="samples per H2D slab in the end-to-end path")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--collective", default="peer", choices=["peer", "nccl"],
                    help="N>1: how the ranks' partial vectors are summed each step")
    ap.add_argument("--no-numa-bind", action="store_true",
                    help="do not pin the process to the CPUs of the GPU's NUMA node (affects the end-to-end host path only)")
    ap.add_argument("--no-overlap", action="store_true",
                    help="launch every step fully serialised after the previous one (no programmatic dependent launch)")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.steps is None:
        args.steps = {"pipeline64": 2000, "pipeline128": 200, "pipeline128x8192": 20, "regdisp512": 200, "fuse2048": 100}[args.workload]
    if args.warmup is None:
        args.warmup = 20
    if args.impl == "reference":
        return run_reference_arm(args)
    return {"pipeline": run_pipeline_arm, "regdisp": run_regdisp_arm, "fuse": run_fuse_arm}[wl["kind"]](args, wl)


if __name__ == "__main__":
    sys.exit(main())
