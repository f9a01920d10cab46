#!/usr/bin/env python
"""Regenerates the `rep` golden fixture: a small DB whose read pairs yield SEVERAL local alignments each, so
that the reference's Handle_Redundancies (dalign/filter.c:1804-2077: Entwine / Fusion / containment / Bridge)
actually changes the output.  The simulator's reads never do that.

    python tests/golden/make_repeats.py        # needs oracle/_ref (built from /root/reference)

Data model (seeded): a random genome with a few tandem-duplicated stretches; ~10 kbp reads at 12 % error
(insert .733 / delete .2 / substitute .067 like db/simulator.c:93-95); half of the reads carry one
"damage" -- mostly a 0.5-1.5 kbp window at 24-34 % error (borderline for -e.70: the forward and the reverse wave
give up at different points inside it), sometimes a junk window or a private insertion of 40-160 random bases --
which breaks an overlap into two alignments that touch, overlap in parallel, or contain one another.  With the
default seed the reference reports 65 containments, 5 fusions and 4 bridges for `daligner R.1 R.1 R.2`.
Outputs (committed): tests/golden/rep/{R.db,.R.idx,.R.bps} and tests/golden/rep/ref/<case>/*.las from the
UNMODIFIED reference daligner."""
import json
import os
import shutil
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = os.path.join(ROOT, "oracle", "_ref")
OUT = os.path.join(ROOT, "tests", "golden", "rep")

CASES = {
    "rep_default": dict(flags=["-j4", "-M0"], blocks=["R.1", "R.1", "R.2"]),
    "rep_b2":      dict(flags=["-j4", "-M0"], blocks=["R.2", "R.2"]),
    "rep_asym":    dict(flags=["-j2", "-M0", "-A"], blocks=["R.2", "R.1"]),
    "rep_s200":    dict(flags=["-j4", "-M0", "-s200", "-e.72"], blocks=["R.1", "R.1"]),
}


def mutate(rng, s, e):
    out = []
    for c in s:
        if rng.random() < e:
            r = rng.random()
            if r < .733:
                out.append(int(rng.integers(4))); out.append(int(c))
            elif r < .933:
                continue
            else:
                out.append(int((c + 1 + rng.integers(3)) % 4))
        else:
            out.append(int(c))
    return np.array(out, np.uint8)


def make_fasta(path, seed=32, genome=60000, cov=27, noisy=(.24, .34), nlen=(500, 1500), fractions=(.03, .50, .55)):
    rng = np.random.default_rng(seed)
    G = rng.integers(0, 4, genome).astype(np.uint8)
    for _ in range(4):                                    # tandem duplications
        p, ln, cp = int(rng.integers(2000, genome - 6000)), int(rng.integers(250, 700)), int(rng.integers(2, 5))
        unit = G[p:p + ln].copy()
        G = np.concatenate([G[:p]] + [mutate(rng, unit, .03) for _ in range(cp)] + [G[p + ln:]])
    genome = len(G)
    total, n = 0, 0
    with open(path, "w") as f:
        while total < cov * genome:
            ln = int(np.clip(rng.normal(9000, 2500), 2500, 16000))
            p = int(rng.integers(0, genome - ln))
            seg = G[p:p + ln].copy()
            kind = rng.random()
            if kind < fractions[0]:                       # junk window
                w = int(rng.integers(120, 500)); q = int(rng.integers(1000, ln - 1000 - w))
                seg[q:q + w] = rng.integers(0, 4, w)
            elif kind < fractions[1]:                     # very noisy window
                w = int(rng.integers(*nlen)); q = int(rng.integers(1000, ln - 1000 - w))
                seg = np.concatenate([seg[:q], mutate(rng, seg[q:q + w], rng.uniform(*noisy)), seg[q + w:]])
            elif kind < fractions[2]:                     # private insertion
                w = int(rng.integers(40, 160)); q = int(rng.integers(1000, ln - 1000))
                seg = np.concatenate([seg[:q], rng.integers(0, 4, w).astype(np.uint8), seg[q:]])
            read = mutate(rng, seg, .12)
            if rng.random() < .5:
                read = (3 - read)[::-1]
            s = "".join("acgt"[c] for c in read)
            n += 1
            f.write(f">Rep/{n}/0_{len(s)} RQ=0.850\n")
            for i in range(0, len(s), 80):
                f.write(s[i:i + 80] + "\n")
            total += len(s)
    return n


def main():
    if os.path.isdir(OUT):
        shutil.rmtree(OUT)
    os.makedirs(OUT)
    fa = os.path.join(OUT, "rep.fasta")
    make_fasta(fa)
    sh = lambda cmd, **kw: subprocess.run(cmd, check=True, **kw)
    sh([os.path.join(REF, "FA2db"), "-x", "1000", "R", "rep.fasta"], cwd=OUT, stdout=subprocess.DEVNULL)
    sh([os.path.join(REF, "DBsplit"), "-s1", "R"], cwd=OUT, stdout=subprocess.DEVNULL)       # 1 Mbp blocks -> 2 blocks
    os.remove(fa)
    for case, c in CASES.items():
        sh([os.path.join(REF, "daligner")] + c["flags"] + c["blocks"], cwd=OUT, stdout=subprocess.DEVNULL)
        d = os.path.join(OUT, "ref", case)
        os.makedirs(d)
        for sub in sorted(os.listdir(OUT)):
            if sub.startswith("d001_"):
                for fn in os.listdir(os.path.join(OUT, sub)):
                    shutil.move(os.path.join(OUT, sub, fn), os.path.join(d, fn))
                shutil.rmtree(os.path.join(OUT, sub))
    with open(os.path.join(OUT, "cases.json"), "w") as f:
        json.dump(CASES, f, indent=1)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
