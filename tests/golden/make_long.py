#!/usr/bin/env python
"""Regenerates the `long` golden fixture: Nanopore-like reads (tens of kbp, 15 % error) from the reference's own
simulator, compared by the UNMODIFIED reference daligner with `-e.65 -l1000` (BASELINE.json configs[3] in miniature).
Reads longer than 50 kbp make report_thread cut a pair's hits into several A-panels (filter.c:2251-2266); at 15 % error
the wave fronts outgrow 32 and 64 diagonals, which takes the parked-pass resume and the strided fallback of the CUDA path.

    python tests/golden/make_long.py

Outputs (committed): tests/golden/long/L.db + hidden files, tests/golden/long/ref/<case>/*.las, cases.json"""
import json
import os
import shutil
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = os.path.join(ROOT, "oracle", "_ref")
OUT = os.path.join(ROOT, "tests", "golden", "long")

CASES = {
    "long_default": dict(flags=["-j4", "-M0", "-e.65", "-l1000"], blocks=["L.1", "L.1"]),
    "long_s126":    dict(flags=["-j2", "-M0", "-e.65", "-l1000", "-s126", "-k12", "-h30"], blocks=["L.1", "L.1"]),
}


def main():
    if os.path.exists(OUT):
        shutil.rmtree(OUT)
    os.makedirs(OUT)
    fa = os.path.join(OUT, "sim.fasta")
    with open(fa, "w") as f:
        subprocess.run([os.path.join(REF, "simulator"), "0.15", "-c5", "-r11", "-e.15", "-m60000", "-s20000", "-x30000"],
                       check=True, stdout=f)
    subprocess.run([os.path.join(REF, "FA2db"), "-x", "1000", "L", "sim.fasta"], cwd=OUT, check=True,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    subprocess.run([os.path.join(REF, "DBsplit"), "-s50", "L"], cwd=OUT, check=True, stdout=subprocess.DEVNULL)
    os.remove(fa)
    for fn in os.listdir(OUT):
        if fn.startswith(".") and fn.split(".")[2] not in ("idx", "bps"):
            os.remove(os.path.join(OUT, fn))
    for name, case in CASES.items():
        subprocess.run([os.path.join(REF, "daligner")] + case["flags"] + case["blocks"], cwd=OUT, check=True,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        dst = os.path.join(OUT, "ref", name)
        os.makedirs(dst)
        for d in sorted(os.listdir(OUT)):
            if d.startswith("d001_"):
                for fn in os.listdir(os.path.join(OUT, d)):
                    shutil.move(os.path.join(OUT, d, fn), os.path.join(dst, fn))
                shutil.rmtree(os.path.join(OUT, d))
        print(name, {fn: os.path.getsize(os.path.join(dst, fn)) for fn in sorted(os.listdir(dst))})
    with open(os.path.join(OUT, "cases.json"), "w") as f:
        json.dump(CASES, f, indent=1)


if __name__ == "__main__":
    main()
