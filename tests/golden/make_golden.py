#!/usr/bin/env python
"""Regenerates the committed golden fixtures from the UNMODIFIED reference tools in oracle/_ref
(built from /root/reference by oracle/Makefile).  Run here, in the build container:

    python tests/golden/make_golden.py

Outputs (all small, committed):
    tests/golden/tiny/         T.db .T.idx .T.bps .T.dust.* .T.rnd.*   -- a 2-block Dazzler DB
    tests/golden/tiny/ref/<case>/*.las                                  -- reference daligner output
    tests/golden/cases.json                                             -- flags of every case

Data: db/simulator.c (seeded), FA2db, DBsplit, DBdust -- exactly the reference's own tools.
A second, synthetic interval track 'rnd' (random intervals) exercises the masked tuple path.
"""
import json
import os
import shutil
import struct
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = os.path.join(ROOT, "oracle", "_ref")
OUT = os.path.join(ROOT, "tests", "golden", "tiny")
sys.path.insert(0, ROOT)

CASES = {
    "default":   dict(flags=["-j4", "-M0"], blocks=["T.1", "T.1", "T.2"]),
    "b2":        dict(flags=["-j2", "-M0"], blocks=["T.2", "T.1", "T.2"]),
    "dust":      dict(flags=["-j4", "-M0", "-mdust"], blocks=["T.1", "T.1", "T.2"]),
    "rndmask":   dict(flags=["-j4", "-M0", "-mrnd"], blocks=["T.1", "T.1", "T.2"]),
    "twomasks":  dict(flags=["-j4", "-M0", "-mrnd", "-mdust"], blocks=["T.2", "T.2"]),
    "identity":  dict(flags=["-j1", "-M0", "-I"], blocks=["T.1", "T.1"]),
    "asym":      dict(flags=["-j4", "-M0", "-A"], blocks=["T.2", "T.1", "T.2"]),
    "k12":       dict(flags=["-j4", "-M0", "-k12", "-w5", "-h30", "-e.75", "-l500", "-s50"], blocks=["T.1", "T.2"]),
    "k16s200":   dict(flags=["-j8", "-M0", "-k16", "-s200", "-t12"], blocks=["T.1", "T.1"]),
    "memlimit":  dict(flags=["-j4", "-M1"], blocks=["T.1", "T.2"]),
    "notrace":   dict(flags=["-j4", "-M0", "-T"], blocks=["T.1", "T.2"]),
    "onlyid":    dict(flags=["-j4", "-M0", "-O"], blocks=["T.1", "T.1", "T.2"]),
}


def sh(cmd, **kw):
    subprocess.run(cmd, check=True, **kw)


def write_random_track(db_dir, name, seed=5):
    """A Dazzler interval track (.anno/.data, 8-byte offsets) with random masked intervals."""
    from damar_b200 import dbio
    blk = dbio.read_block(os.path.join(db_dir, "T"))
    rng = np.random.default_rng(seed)
    anno = [0]
    data = []
    for i in range(blk.nreads):
        n = int(blk.rlen[i])
        pos, iv = 0, []
        while True:
            pos += int(rng.integers(5, 1500))
            ln = int(rng.integers(1, 400))
            if pos + ln >= n:
                break
            iv += [pos, pos + ln]
            pos += ln + int(rng.integers(0, 3) == 0) * int(rng.integers(1, 30))
        if rng.integers(0, 10) == 0 and n > 50:      # sometimes mask the read start / end exactly
            iv = [0, 20] + [x for x in iv if x > 25]
            if len(iv) % 2:
                iv = iv[:-1]
        data += iv
        anno.append(4 * len(data))
    with open(os.path.join(db_dir, f".T.{name}.anno"), "wb") as f:
        f.write(struct.pack("<ii", blk.nreads, 8))
        f.write(np.asarray(anno, dtype="<i8").tobytes())
    with open(os.path.join(db_dir, f".T.{name}.data"), "wb") as f:
        f.write(np.asarray(data, dtype="<i4").tobytes())


def main():
    if os.path.exists(OUT):
        shutil.rmtree(OUT)
    os.makedirs(OUT)
    with open(os.path.join(OUT, "sim.fasta"), "w") as f:
        sh([os.path.join(REF, "simulator"), "0.12", "-c16", "-r11", "-e.12", "-m3000", "-s900", "-x800"], stdout=f)
    sh([os.path.join(REF, "FA2db"), "-x", "500", "T", "sim.fasta"], cwd=OUT)
    sh([os.path.join(REF, "DBsplit"), "-s1", "T"], cwd=OUT)
    sh([os.path.join(REF, "DBdust"), "T"], cwd=OUT)
    os.remove(os.path.join(OUT, "sim.fasta"))
    for junk in (".T.pacbio.anno", ".T.pacbio.data", ".T.seqID.anno", ".T.seqID.data"):
        p = os.path.join(OUT, junk)
        if os.path.exists(p):
            os.remove(p)
    write_random_track(OUT, "rnd")
    print(open(os.path.join(OUT, "T.db")).read())
    for name, case in CASES.items():
        for d in os.listdir(OUT):
            if d.startswith("d0"):
                shutil.rmtree(os.path.join(OUT, d))
        sh([os.path.join(REF, "daligner")] + case["flags"] + case["blocks"], cwd=OUT, stdout=subprocess.DEVNULL)
        dst = os.path.join(OUT, "ref", name)
        os.makedirs(dst)
        n = 0
        for d in sorted(os.listdir(OUT)):
            if d.startswith("d0"):
                for fn in os.listdir(os.path.join(OUT, d)):
                    shutil.move(os.path.join(OUT, d, fn), os.path.join(dst, fn))
                    n += os.path.getsize(os.path.join(dst, fn))
                shutil.rmtree(os.path.join(OUT, d))
        print(f"{name:10s} {case['flags']} {case['blocks']} -> {n} bytes")
    with open(os.path.join(ROOT, "tests", "golden", "cases.json"), "w") as f:
        json.dump(CASES, f, indent=1)


if __name__ == "__main__":
    main()
