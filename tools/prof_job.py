#!/usr/bin/env python
"""Runs ONE daligner job of a bench workload on the GPU (for ncu captures). usage: prof_job.py <workload> <a> [b...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from damar_b200 import api, dbio  # noqa: E402

wl = sys.argv[1]
a = int(sys.argv[2])
bs = [int(x) for x in sys.argv[3:]] or [a]
dbdir, root, nblocks = bench.make_db(wl)
api.init()
f = bench.DAL_FLAGS
s = api.Settings(kmer=f["k"], binshift=f["w"], hitmin=f["h"], ave_corr=f["e"], minlen=f["l"], trace_spacing=f["s"], mem_limit_gb=f["M"])
blocks = {i: dbio.read_block(os.path.join(dbdir, f"{root}.{i}")) for i in set([a] + bs)}
out = api.daligner(blocks[a], [blocks[b] for b in bs], s, as_bytes=False)
print({k: v.novl for k, v in out.items()}, api.timings())
