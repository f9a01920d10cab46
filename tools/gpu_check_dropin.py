#!/usr/bin/env python
"""The drop-in boundary at scale: the UNMODIFIED reference host (daligner.c + align.c + DB.c) linked with libdamar_b200.so
through the filter.h shim (oracle/_ref/daligner_b200) against the unmodified reference binary on bench workload `small`
(2 x 10 Mbp), several flag sets; every .las byte-compared (padding masked).  Test infrastructure (executes oracle/_ref)."""
import os
import shutil
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from damar_b200 import dbio  # noqa: E402

FLAG_SETS = [[], ["-t12"], ["-I"], ["-A"], ["-s200"], ["-k16", "-h50"], ["-b"], ["-e.85"]]


def run(exe, flags, dbdir, root):
    for d in os.listdir(dbdir):
        if d.startswith("d001_"):
            shutil.rmtree(os.path.join(dbdir, d))
    t0 = time.time()
    subprocess.run([os.path.join(ROOT, "oracle", "_ref", exe), "-j4", "-M32"] + flags + [f"{root}.2", f"{root}.1", f"{root}.2"],
                   cwd=dbdir, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    dt = time.time() - t0
    out = {}
    for d in sorted(os.listdir(dbdir)):
        if d.startswith("d001_"):
            for fn in os.listdir(os.path.join(dbdir, d)):
                out[fn] = open(os.path.join(dbdir, d, fn), "rb").read()
    return out, dt


def main():
    dbdir, root, _ = bench.make_db("small")
    allok = True
    for fl in FLAG_SETS:
        ref, t_ref = run("daligner", fl, dbdir, root)
        got, t_got = run("daligner_b200", fl, dbdir, root)
        ok = sorted(ref) == sorted(got) and all(dbio.las_equal(ref[k], got[k])[0] for k in ref)
        allok &= ok
        n = sum(int.from_bytes(v[:8], "little") for v in got.values())
        print(f"  daligner_b200 -j4 -M32 {' '.join(fl):14s} {len(got)} files {n:7d} overlaps  {t_got:5.2f} s (process start + CUDA init included)"
              f"  reference {t_ref:5.2f} s: {'OK' if ok else 'DIFF'}", flush=True)
    print("PARITY", "PASS" if allok else "FAIL")
    sys.exit(0 if allok else 1)


if __name__ == "__main__":
    main()
