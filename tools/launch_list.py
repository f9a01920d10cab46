#!/usr/bin/env python
"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: total ms, share, launches.
usage: launch_list.py <launches.csv> [header comment ...]"""
import csv
import re
import sys
from collections import defaultdict

rows = list(csv.reader(l for l in open(sys.argv[1], errors="replace") if l.startswith('"')))
h = rows[0]
ki, mi, vi, ui = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("Metric Unit")
tot = defaultdict(float)
cnt = defaultdict(int)
for r in rows[1:]:
    if len(r) <= vi or r[mi] != "gpu__time_duration.sum":
        continue
    v = float(r[vi].replace(",", ""))
    u = r[ui]
    ms = v / 1e6 if u in ("ns", "nsecond") else v / 1e3 if u in ("us", "usecond") else v if u in ("ms", "msecond") else v * 1e3
    name = re.sub(r"\(.*", "", r[ki])
    name = re.sub(r"^void ", "", name).replace("damar::", "").replace("<unnamed>::", "").replace("(anonymous namespace)::", "").replace("unnamed>::", "")
    tot[name] += ms
    cnt[name] += 1
all_ms = sum(tot.values())
for c in sys.argv[2:]:
    print("# " + c)
print(f"# total kernel time {all_ms:.1f} ms over {sum(cnt.values())} launches")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"{v:10.2f} ms {100 * v / all_ms:5.1f}%  n={cnt[k]:5d}  {k}")
