#!/usr/bin/env python
"""Attribute executed warp-instructions of one kernel in an .ncu-rep to source lines.
Usage: ncu_lines.py <rep> <nvdisasm -g -c output of the same build> <mangled kernel name> [top]"""
import re, csv, collections, subprocess, io, sys, glob, os
rep, dis, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
lines = open(dis).read().split('\n')
start = [i for i, l in enumerate(lines) if ('.text.' + kname) in l and l.startswith('//---')][0]
end = next(i for i in range(start + 1, len(lines)) if lines[i].startswith('//---------------------'))
cur = None; off2line = {}
for l in lines[start:end]:
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m:
        off2line[int(m.group(1), 16)] = (cur, m.group(2))
sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
blocks = []; cur = None
for r in csv.reader(io.StringIO(sass)):
    if r and r[0] == "Kernel Name": cur = {'rows': []}; blocks.append(cur); continue
    if r and r[0] == "Address": cur['hdr'] = r; continue
    if cur is not None and len(r) > 5: cur['rows'].append(r)
b = blocks[int(os.environ.get("NCU_BLOCK", "0"))]; h = b['hdr']; iA = h.index('Address'); iI = h.index('Instructions Executed'); iS = h.index('# Samples')
base = int(b['rows'][0][iA], 16)
agg = collections.Counter(); smp = collections.Counter(); ops = collections.defaultdict(collections.Counter)
for r in b['rows']:
    off = int(r[iA], 16) - base; n = int(r[iI] or 0)
    key, txt = off2line.get(off, (None, ''))
    agg[key] += n; smp[key] += int(r[iS] or 0)
    m = re.match(r'\s*(@!?U?P\w+\s+)?([A-Z0-9_.]+)', txt or '')
    ops[key][(m.group(2).split('.')[0] if m else '?')] += n
tot = sum(agg.values()); src = {}
print(f"total warp-instructions {tot}")
root = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
for (k, c) in agg.most_common(top):
    if k is None: print('unknown', c); continue
    f, ln = k
    if f not in src:
        cand = glob.glob(os.path.join(root, '*', 'csrc', f))
        src[f] = open(cand[0]).read().split('\n') if cand else None
    text = src[f][ln - 1].strip()[:70] if src[f] else ''
    mix = ' '.join(f"{o}:{n*100//max(c,1)}" for o, n in ops[k].most_common(3))
    print(f"{f[:24]:24s} L{ln:4d} {c:9d} {100*c/tot:5.1f}% smp={smp[k]:5d} [{mix:34s}] {text}")
