#!/usr/bin/env python
"""Correlate an ncu SASS source page (csv) with nvdisasm -g line info: per source line instruction
counts and stall samples.  usage: ncu_lines.py <src.csv> <nvdisasm.sass> <mangled-kernel-substring> [top]"""
import csv
import re
import sys
from collections import defaultdict

srccsv, sass, kname = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 50
rows = list(csv.reader(open(srccsv)))
hi = [i for i, r in enumerate(rows) if 'Instructions Executed' in r][0]
hdr = rows[hi]
ci, cs, so = hdr.index('Instructions Executed'), hdr.index('# Samples'), hdr.index('Source')
insts = []
for r in rows[hi + 1:]:
    if len(r) > ci:
        try:
            insts.append((int(float(r[ci] or 0)), int(float(r[cs] or 0)), r[so].strip()))
        except ValueError:
            pass
# nvdisasm: walk the function, tracking the current line
lines = open(sass).read().split('\n')
start = [i for i, l in enumerate(lines) if l.startswith('.text.') and kname in l][0]
cur = ('?', 0)
linemap = []
for l in lines[start + 1:]:
    if l.startswith('//-----') and '.text.' in l:
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    if re.match(r'\s+/\*[0-9a-f]{4}\*/', l):
        linemap.append(cur)
print(f"ncu insts {len(insts)}  nvdisasm insts {len(linemap)}", file=sys.stderr)
n = min(len(insts), len(linemap))
agg = defaultdict(lambda: [0, 0, 0])
tot = sum(x[0] for x in insts)
tots = sum(x[1] for x in insts)
for i in range(n):
    a = agg[linemap[i]]
    a[0] += insts[i][0]; a[1] += insts[i][1]; a[2] += 1
src = {}
for (f, ln), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    if f not in src:
        try:
            src[f] = open('/root/repo/damar_b200/csrc/' + f).read().split('\n')
        except OSError:
            src[f] = []
    text = src[f][ln - 1].strip()[:100] if 0 < ln <= len(src[f]) else ''
    print(f"{v[0] / tot * 100:5.1f}% inst {v[1] / max(tots, 1) * 100:5.1f}% stall  sass={v[2]:3d}  {f}:{ln:4d}  {text}")
