#!/usr/bin/env python
"""Times the device radix sort alone (damar_sort_selftest: random keys, checked on the device) at the sizes of the
hot path: packed k-mer records (keys-only, 28 sort bits), (code, gpos) pairs, seed-hit keys.  One line per case:
n, bits, payload, ms, G keys/s, real GB/s (hist read + passes) and contract GB/s (16-B records, 8-bit digits)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from damar_b200 import api  # noqa: E402

CASES = [("kmers packed k=14", 150_000_000, 28, 56, False), ("kmers pairs k=14", 150_000_000, 0, 28, True),
         ("kmers pairs k=20", 150_000_000, 0, 40, True), ("hits packed 44b", 100_000_000, 16, 60, False),
         ("hits pairs 44b", 100_000_000, 0, 44, True), ("hits packed 44b 900M", 900_000_000, 16, 60, False),
         ("small 1M pairs 24b", 1_000_000, 0, 24, True)]
if len(sys.argv) > 1:
    CASES = [c for c in CASES if any(a in c[0] for a in sys.argv[1:])]
maxb = int(os.environ.get("DAMAR_SORT_BITS", "10"))
for name, n, lo, hi, vals in CASES:
    try:
        ms = api.sort_selftest(n, lo, hi, vals, reps=3, seed=7)
    except RuntimeError as e:
        print(f"{name:24s} FAILED: {e}")
        continue
    nbits = hi - lo
    npass = -(-nbits // maxb)
    rec = 12 if vals else 8
    real = n * (8 + npass * 2 * rec) / ms / 1e6
    contract = n * 16 * (1 + 2 * -(-nbits // 8)) / ms / 1e6
    print(f"{name:24s} n={n:>11d} bits={nbits:2d} passes={npass} {ms:8.3f} ms  {n / ms / 1e6:7.2f} Gkeys/s  real {real:7.0f} GB/s  contract {contract:7.0f} GB/s", flush=True)
