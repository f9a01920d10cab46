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
masks = None
if wl not in bench.WORKLOADS:                      # the reduced configs[3] / configs[4] of tools/gpu_check_configs.py
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import gpu_check_configs as G
    cfg = G.CONFIGS[wl]
    bench.WORKLOADS[wl] = dict(sim=cfg["sim"], split=cfg["split"], root=cfg["root"], desc=wl)
    masks = cfg["masks"] or None
dbdir, root, nblocks = bench.make_db(wl)
api.init()
if masks:
    import subprocess
    if not os.path.exists(os.path.join(dbdir, f".{root}.dust.anno")):
        subprocess.run([os.path.join(ROOT, "oracle", "_ref", "DBdust"), root], cwd=dbdir, check=True, stdout=subprocess.DEVNULL)
    s = G.settings(cfg["flags"])
elif wl in ("c4s", "c5s"):
    s = G.settings(cfg["flags"])
else:
    f = bench.DAL_FLAGS
    s = api.Settings(kmer=f["k"], binshift=f["w"], hitmin=f["h"], ave_corr=f["e"], minlen=f["l"], trace_spacing=f["s"], mem_limit_gb=f["M"])
blocks = {i: dbio.read_block(os.path.join(dbdir, f"{root}.{i}"), masks) for i in set([a] + bs)}
out = api.daligner(blocks[a], [blocks[b] for b in bs], s, as_bytes=False)
print({k: v.novl for k, v in out.items()}, api.timings())
