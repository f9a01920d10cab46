#!/usr/bin/env python
"""Regenerates the `edge` golden fixture: two tiny DBs and the reference daligner's answers for the corner cases of the
path -- reads of exactly k and k+1 bases, blocks without any k-mer match, nothing passing -h, and a comparison across
two DBs (asymmetric: only the A side is reported, daligner.c:911-946) -- from the UNMODIFIED reference tools in oracle/_ref.

    python tests/golden/make_edge.py

Outputs (committed): tests/golden/edge/{E,F}.db + hidden files, tests/golden/edge/ref/<case>/*.las, cases.json"""
import json
import os
import random
import shutil
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = os.path.join(ROOT, "oracle", "_ref")
OUT = os.path.join(ROOT, "tests", "golden", "edge")

CASES = {
    "edge_self":  dict(flags=["-j4", "-M0", "-l500"], blocks=["E.1", "E.1"]),
    "edge_none":  dict(flags=["-j4", "-M0", "-l500"], blocks=["F.1", "F.1"]),
    "edge_h":     dict(flags=["-j4", "-M0", "-l500", "-h2000"], blocks=["E.1", "E.1"]),
    "edge_cross": dict(flags=["-j4", "-M0", "-l500"], blocks=["F.1", "E.1", "F.1"]),
    "edge_k16":   dict(flags=["-j2", "-M0", "-l500", "-k14", "-t3", "-I"], blocks=["E.1", "E.1"]),
}


def main():
    rng = random.Random(11)
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

    if os.path.exists(OUT):
        shutil.rmtree(OUT)
    os.makedirs(OUT)
    g = rnd(9000)
    dbs = {"e": [mut(g[0:4000], .08), rnd(14), mut(g[2000:6500], .08), rnd(15), mut(g[4500:9000], .08), rnd(2500),
                 mut(g[1000:8000], .08), rnd(40)],
           "f": [rnd(3000), rnd(2800), rnd(14), rnd(3500)]}
    for name, reads in dbs.items():
        fa = os.path.join(OUT, name + ".fasta")
        with open(fa, "w") as f:
            for i, r in enumerate(reads):
                f.write(f">{name}/{i + 1}/0_{len(r)} RQ=0.85\n")
                for j in range(0, len(r), 80):
                    f.write(r[j:j + 80] + "\n")
        subprocess.run([os.path.join(REF, "FA2db"), "-x", "5", name.upper(), name + ".fasta"], cwd=OUT, check=True,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        subprocess.run([os.path.join(REF, "DBsplit"), "-s1", name.upper()], cwd=OUT, check=True, stdout=subprocess.DEVNULL)
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
