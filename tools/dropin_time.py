#!/usr/bin/env python
"""Wall time of the HPCdaligner plan of a bench workload run by the DROP-IN host (the unmodified dalign/daligner.c linked
against libdamar_b200.so through the filter.h shim, oracle/_ref/daligner_b200: one process per job, as the reference's
plan would run it) beside the reference binary, with a byte comparison of every .las file.  usage: dropin_time.py [workload]"""
import os
import shutil
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from damar_b200 import api  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
dbdir, root, nblocks = bench.make_db(wl)
P = bench.pow2_floor(os.cpu_count() or 1)


def run(exe, tag):
    for d in os.listdir(dbdir):
        if d.startswith("d001_"):
            shutil.rmtree(os.path.join(dbdir, d))
    t0 = time.perf_counter()
    per = []
    for a, bs in bench.plan(nblocks):
        t1 = time.perf_counter()
        subprocess.run([exe] + bench.ref_flags(P) + [f"{root}.{a}"] + [f"{root}.{b}" for b in bs], cwd=dbdir, check=True,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, env={**os.environ, "CUDA_VISIBLE_DEVICES": "0"})
        per.append(round(time.perf_counter() - t1, 2))
    dt = time.perf_counter() - t0
    files = {}
    for d in sorted(os.listdir(dbdir)):
        if d.startswith("d001_"):
            for fn in os.listdir(os.path.join(dbdir, d)):
                files[fn] = open(os.path.join(dbdir, d, fn), "rb").read()
            shutil.rmtree(os.path.join(dbdir, d))
    novl = sum(int.from_bytes(v[:8], "little") for v in files.values())
    print(f"{tag}: {dt:.2f} s for {len(per)} jobs {per}, {novl} overlaps, {novl / dt:.0f} overlaps/s", flush=True)
    return files, dt


ours, t_ours = run(os.path.join(ROOT, "oracle", "_ref", "daligner_b200"), "drop-in host (daligner.c + shim + libdamar_b200.so)")
ref, t_ref = run(os.path.join(ROOT, "oracle", "_ref", "daligner"), f"reference daligner -j{P}")
bad = [fn for fn in ref if fn not in ours or api.las_compare(ref[fn], ours[fn]) != -1]
print(f"files {len(ref)}, byte-identical {len(ref) - len(bad)}, speed-up {t_ref / t_ours:.1f}x", "MISMATCH " + str(bad) if bad else "")
sys.exit(1 if bad else 0)
