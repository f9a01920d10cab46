#!/usr/bin/env python
"""Text timeline of a DAMAR_TRACE file (lines "thread scope start_ms end_ms" of the library's profiled host scopes):
one row per host thread, one character per `res` ms = the innermost scope that covered most of it.
usage: trace_view.py <file> [res_ms] [from_ms] [to_ms]"""
import sys
from collections import Counter, defaultdict

LET = {"wait": "w", "cudaMalloc": "M", "ra.align_run": "A", "ra.worklist": "L", "ra.fuse": "F", "ra.download": "d", "sh.count": "c",
       "sh.merge": "m", "sh.sort+free": "s", "sort_kmers": "K", "block_upload": "U", "block_from_view": "V", "dsink_serialise": "D",
       "mf.seed_hits": "S", "mf.report_and_align": "R", "sink_add": "a", "sink_serialise": "z"}
recs = []
for line in open(sys.argv[1]):
    t, name, a, b = line.split()
    recs.append((int(t), name, float(a), float(b)))
res = float(sys.argv[2]) if len(sys.argv) > 2 else 5.0
org = min(r[2] for r in recs)
lo = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
hi = float(sys.argv[4]) if len(sys.argv) > 4 else max(r[3] for r in recs) - org
fine = 0.25                                            # ms per painted cell
n = int((hi - lo) / fine) + 1
rows = defaultdict(lambda: [" "] * n)
for t, name, a, b in sorted(recs, key=lambda r: -(r[3] - r[2])):      # parents first, children paint over them
    a, b = a - org, b - org
    if b < lo or a > hi:
        continue
    i0, i1 = max(0, int((a - lo) / fine)), min(n, int((b - lo) / fine) + 1)
    row = rows[t]
    c = LET.get(name, "?")
    for i in range(i0, i1):
        row[i] = c
per = int(res / fine)
print("legend:", " ".join(f"{v}={k}" for k, v in LET.items()), "' '=outside the library")
print(f"window {lo:.0f}..{hi:.0f} ms, {res} ms per character")
tot = Counter()
for t in sorted(rows):
    row = rows[t]
    tot.update(row)
    print(f"T{t:02d} " + "".join(Counter(row[i:i + per]).most_common(1)[0][0] for i in range(0, n, per)))
all_cells = sum(tot.values())
print("share of thread time:", ", ".join(f"{k!r} {100 * v / all_cells:.1f}%" for k, v in tot.most_common()))
