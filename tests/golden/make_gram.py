#!/usr/bin/env python
"""Regenerates the `gram` golden fixture: k-mer groups above the MAXGRAM cap, oversized equal-code runs and a read pair
with more than 127 local alignments -- cases the simulator-made fixtures are too small to contain -- answered by the
UNMODIFIED reference tools in oracle/_ref.

  R: 170 reads that share one exact 400 bp element (every k-mer of it has 170*169/2 = 14365 >= 10000 mutual matches in
     a self comparison: dropped by the default -M = physical memory, kept with -M0, filter.c:2634-2663), and two reads
     with a 3000 bp poly-A stretch (one 14-mer run of ~6000 entries per block).
  W: two reads that carry 12 copies each of one 350 bp element between random spacers of different lengths: 144 local
     alignments on 144 different diagonals for ONE read pair (the reference keeps them in a realloc'ed list,
     filter.c:2318-2380), 288 records with -l100.

    python tests/golden/make_gram.py

Outputs (committed): tests/golden/gram/{R,W}.db + hidden files, tests/golden/gram/ref/<case>/*.las, cases.json"""
import json
import os
import random
import shutil
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = os.path.join(ROOT, "oracle", "_ref")
OUT = os.path.join(ROOT, "tests", "golden", "gram")

CASES = {
    "gram_default": dict(flags=["-j4", "-l150"], blocks=["R.1", "R.1"]),            # -M = physical RAM: cap at MAXGRAM
    "gram_M0":      dict(flags=["-j4", "-M0", "-l150"], blocks=["R.1", "R.1"]),     # no cap: the shared element aligns
    "gram_M1":      dict(flags=["-j4", "-M1", "-l150"], blocks=["R.1", "R.1"]),
    "many_la":      dict(flags=["-j4", "-M0", "-l100"], blocks=["W.1", "W.1"]),
    "many_la_A":    dict(flags=["-j2", "-M0", "-l100", "-A"], blocks=["W.1", "W.1"]),
}


def main():
    rng = random.Random(5)
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
    elem = rnd(400)
    rreads = [rnd(rng.randint(300, 700)) + elem + rnd(rng.randint(200, 500)) for _ in range(170)]
    g = rnd(6000)
    rreads += [mut(g[0:3500], .06), mut(g[1500:5000], .06), mut(g[2500:6000], .06)]      # ordinary overlaps beside the repeats
    rreads += [rnd(400) + "a" * 3000 + rnd(500), rnd(600) + "a" * 3000 + rnd(300)]
    f350 = rnd(350)
    wreads = []
    for _ in range(2):
        parts = [rnd(rng.randint(60, 400))]
        for _ in range(12):
            parts += [mut(f350, .03), rnd(rng.randint(60, 400))]
        wreads.append("".join(parts))
    wreads += [mut(g[500:4000], .08), mut(g[2000:5500], .08)]
    for name, reads in (("r", rreads), ("w", wreads)):
        fa = os.path.join(OUT, name + ".fasta")
        with open(fa, "w") as f:
            for i, r in enumerate(reads):
                f.write(f">{name}/{i + 1}/0_{len(r)} RQ=0.85\n")
                for j in range(0, len(r), 80):
                    f.write(r[j:j + 80] + "\n")
        subprocess.run([os.path.join(REF, "FA2db"), "-x", "100", name.upper(), name + ".fasta"], cwd=OUT, check=True,
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
