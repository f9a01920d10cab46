#!/usr/bin/env python
"""Reads a DAMAR_WAVE_LOG file (u64 quadruples {start, first idle lane, end, SM | wide << 16 | grid << 32} per warp of every
wave_tasks launch) and prints how the SMs were occupied: warp residency against the capacity (148 SMs x 12 warps), the
share of it spent after the task queue ran dry (lanes waiting for the longest passes), per time bin.
usage: wave_log.py <file> [bin_ms] [warps_per_sm]"""
import sys
import numpy as np

a = np.fromfile(sys.argv[1], dtype=np.uint64).reshape(-1, 4)
bin_ms = float(sys.argv[2]) if len(sys.argv) > 2 else 25.0
wps = int(sys.argv[3]) if len(sys.argv) > 3 else 12
t0, td, t1 = a[:, 0].astype(np.int64), a[:, 1].astype(np.int64), a[:, 2].astype(np.int64)
wide = ((a[:, 3] >> np.uint64(16)) & np.uint64(1)).astype(bool)
grid = (a[:, 3] >> np.uint64(32)).astype(np.int64)
td = np.where(a[:, 1] == np.uint64(0xFFFFFFFFFFFFFFFF), t1, td)
org = t0.min()
t0, td, t1 = (t0 - org) / 1e6, (td - org) / 1e6, (t1 - org) / 1e6
span = t1.max()
print(f"{len(a)} warp records, span {span:.1f} ms, narrow {np.sum(~wide)}, wide {np.sum(wide)}")
res = t1 - t0
print(f"warp residency: narrow {res[~wide].sum() / 1e3:.2f} warp-s (busy until the queue ran dry {np.sum((td - t0)[~wide]) / 1e3:.2f}, "
      f"tail {np.sum((t1 - td)[~wide]) / 1e3:.2f}), wide {res[wide].sum() / 1e3:.2f} warp-s; capacity {148 * wps * span / 1e3:.2f} warp-s")
nb = int(span / bin_ms) + 1
occ = np.zeros(nb); occ_bulk = np.zeros(nb); occ_wide = np.zeros(nb)


def add(dst, s, e):
    for b in range(nb):
        lo, hi = b * bin_ms, (b + 1) * bin_ms
        dst[b] += np.clip(np.minimum(e, hi) - np.maximum(s, lo), 0, None).sum()


add(occ, t0[~wide], t1[~wide]); add(occ_bulk, t0[~wide], td[~wide]); add(occ_wide, t0[wide], t1[wide])
cap = 148 * wps * bin_ms
print("bin start ms: narrow warps resident / of which before the queue ran dry / wide warps (fractions of 148 x %d warps)" % wps)
for b in range(nb):
    print(f"{b * bin_ms:8.0f}  {occ[b] / cap:5.2f}  {occ_bulk[b] / cap:5.2f}  {occ_wide[b] / cap:5.2f}")
