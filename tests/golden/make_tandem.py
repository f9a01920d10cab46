#!/usr/bin/env python
"""Regenerates the `tan` golden fixture for datander (scrub/datander.c + scrub/tandem.c): reads with tandem arrays
(units of 40-400 bp, 4-40 copies, 2-10 % divergence between copies) between random sequence, one of them longer than a
50 kbp panel, answered by the UNMODIFIED reference datander built by oracle/Makefile.

    python tests/golden/make_tandem.py

Outputs (committed): tests/golden/tan/N.db + hidden files, tests/golden/tan/ref/<case>/*.las, cases.json"""
import filecmp
import json
import os
import random
import shutil
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = os.path.join(ROOT, "oracle", "_ref")
OUT = os.path.join(ROOT, "tests", "golden", "tan")

CASES = {
    "tan_default": dict(flags=["-j4"], blocks=["N.1"]),
    "tan_j1":      dict(flags=["-j1"], blocks=["N.1"]),
    "tan_l200":    dict(flags=["-j2", "-l200", "-s50"], blocks=["N.1"]),
    "tan_k14":     dict(flags=["-j4", "-k14", "-w5", "-h40", "-e.8", "-s126"], blocks=["N.2"]),
    "tan_b2":      dict(flags=["-j8"], blocks=["N.2"]),
}


def main():
    rng = random.Random(23)
    rnd = lambda n: "".join(rng.choice("acgt") for _ in range(n))

    def mut(s, e):
        o = []
        for c in s:
            r = rng.random()
            if r < e * 0.3:
                continue
            if r < e * 0.6:
                o.append(rng.choice("acgt"))
            o.append(c if rng.random() > e * 0.4 else rng.choice("acgt"))
        return "".join(o)

    def array(unit_len, copies, div):
        u = rnd(unit_len)
        return "".join(mut(u, div) for _ in range(copies))

    reads = []
    for i in range(230):
        parts = [rnd(rng.randint(200, 2500))]
        for _ in range(rng.randint(0, 3)):
            parts += [array(rng.randint(40, 400), rng.randint(4, 40), rng.choice([.02, .05, .10])), rnd(rng.randint(100, 3000))]
        reads.append("".join(parts))
    # one read longer than a panel (PANEL_SIZE 50000, tandem.c:73), arrays on both sides of the panel boundary
    reads.insert(20, rnd(20000) + array(180, 60, .06) + rnd(16000) + array(90, 150, .08) + rnd(9000) + array(350, 12, .04) + rnd(3000))
    reads.append("ac" * 900 + rnd(700) + "a" * 800 + rnd(300))                   # low-complexity stretches
    if os.path.exists(OUT):
        shutil.rmtree(OUT)
    os.makedirs(OUT)
    fa = os.path.join(OUT, "n.fasta")
    with open(fa, "w") as f:
        for i, r in enumerate(reads):
            f.write(f">n/{i + 1}/0_{len(r)} RQ=0.85\n")
            for j in range(0, len(r), 80):
                f.write(r[j:j + 80] + "\n")
    subprocess.run([os.path.join(REF, "FA2db"), "-x", "100", "N", "n.fasta"], cwd=OUT, check=True,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    subprocess.run([os.path.join(REF, "DBsplit"), "-s1", "N"], cwd=OUT, check=True, stdout=subprocess.DEVNULL)
    os.remove(fa)
    for fn in os.listdir(OUT):
        if fn.startswith(".") and fn.split(".")[2] not in ("idx", "bps"):
            os.remove(os.path.join(OUT, fn))
    print(open(os.path.join(OUT, "N.db")).read())
    for name, case in CASES.items():
        subprocess.run([os.path.join(REF, "datander")] + case["flags"] + ["-o", "tan"] + case["blocks"], cwd=OUT, check=True,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        dst = os.path.join(OUT, "ref", name)
        os.makedirs(dst)
        for fn in os.listdir(os.path.join(OUT, "tan")):
            shutil.move(os.path.join(OUT, "tan", fn), os.path.join(dst, fn))
        shutil.rmtree(os.path.join(OUT, "tan"))
        print(name, {fn: os.path.getsize(os.path.join(dst, fn)) for fn in sorted(os.listdir(dst))})
    assert filecmp.cmp(os.path.join(OUT, "ref", "tan_default", "N.1.N.1.las"), os.path.join(OUT, "ref", "tan_j1", "N.1.N.1.las"), shallow=False)
    with open(os.path.join(OUT, "cases.json"), "w") as f:
        json.dump(CASES, f, indent=1)


if __name__ == "__main__":
    main()
