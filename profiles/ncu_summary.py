#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page + SASS opcode histogram) -> stdout.  Usage: ncu_summary.py file.ncu-rep"""
import collections, csv, io, re, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
WANT = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.sum",
        "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_drain_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_imc_miss_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio",
        ]
for w in WANT:
    if w in hdr:
        i = hdr.index(w)
        print(f"{w:82s} {units[i]:16s} {[r[i][:48] for r in rows[2:]]}")
sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
cur = None
blocks = []
for r in csv.reader(io.StringIO(sass)):
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}; blocks.append(cur); continue
    if r and r[0] == "Address":
        cur["hdr"] = r; continue
    if cur is not None and len(r) > 5:
        cur["rows"].append(r)
if blocks:
    b = blocks[0]; h = b["hdr"]
    iA, iI, iS = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
    tot = sum(int(r[iI] or 0) for r in b["rows"])
    print(f"\nSASS of {b['name'][:80]}: {len(b['rows'])} instructions, {tot} warp-instructions executed")
    op, smp = collections.Counter(), collections.Counter()
    for r in b["rows"]:
        m = re.match(r"\s*(@!?U?P\w+\s+)?([A-Z0-9_.]+)", r[iA])
        o = m.group(2).split(".")[0] if m else "?"
        op[o] += int(r[iI] or 0); smp[o] += int(r[iS] or 0)
    for o, c in op.most_common(24):
        print(f"  {o:10s} {c:10d} {100*c/tot:5.1f}%   stall samples {smp[o]}")
