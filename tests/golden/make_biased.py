#!/usr/bin/env python
"""Regenerates the `bias` golden fixture: an AT-rich DB (simulator -b.7) for daligner's -b option (variable-length
k-mers by log-frequency, dalign/filter.c:549-688), from the UNMODIFIED reference tools in oracle/_ref.

    python tests/golden/make_biased.py

Outputs (committed): tests/golden/bias/{B.db,.B.idx,.B.bps,.B.dust.*}, tests/golden/bias/ref/<case>/*.las, cases.json"""
import json
import os
import shutil
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = os.path.join(ROOT, "oracle", "_ref")
OUT = os.path.join(ROOT, "tests", "golden", "bias")

CASES = {
    "bias_b":      dict(flags=["-j4", "-M0", "-b"], blocks=["B.1", "B.1", "B.2"]),
    "bias_b_dust": dict(flags=["-j2", "-M0", "-b", "-mdust"], blocks=["B.2", "B.1", "B.2"]),
    "bias_b_k16":  dict(flags=["-j4", "-M0", "-b", "-k16", "-h40", "-t30"], blocks=["B.1", "B.1"]),
    "bias_b_h90":  dict(flags=["-j4", "-M0", "-b", "-h90"], blocks=["B.1", "B.2"]),
    "bias_plain":  dict(flags=["-j4", "-M0"], blocks=["B.1", "B.2"]),
    "bias_plain_h90": dict(flags=["-j4", "-M0", "-h90"], blocks=["B.1", "B.2"]),
}


def sh(cmd, **kw):
    subprocess.run(cmd, check=True, **kw)


def main():
    if os.path.exists(OUT):
        shutil.rmtree(OUT)
    os.makedirs(OUT)
    with open(os.path.join(OUT, "sim.fasta"), "w") as f:
        sh([os.path.join(REF, "simulator"), "0.10", "-c14", "-b.7", "-r23", "-e.12", "-m3000", "-s900", "-x800"], stdout=f)
    sh([os.path.join(REF, "FA2db"), "-x", "500", "B", "sim.fasta"], cwd=OUT, stdout=subprocess.DEVNULL)
    sh([os.path.join(REF, "DBsplit"), "-s1", "B"], cwd=OUT, stdout=subprocess.DEVNULL)
    sh([os.path.join(REF, "DBdust"), "B"], cwd=OUT, stdout=subprocess.DEVNULL)
    os.remove(os.path.join(OUT, "sim.fasta"))
    for fn in os.listdir(OUT):
        if fn.startswith(".B.") and fn.split(".")[2] not in ("idx", "bps", "dust"):
            os.remove(os.path.join(OUT, fn))
    for name, case in CASES.items():
        sh([os.path.join(REF, "daligner")] + case["flags"] + case["blocks"], cwd=OUT, stdout=subprocess.DEVNULL)
        dst = os.path.join(OUT, "ref", name)
        os.makedirs(dst)
        for d in sorted(os.listdir(OUT)):
            if d.startswith("d001_"):
                for fn in os.listdir(os.path.join(OUT, d)):
                    shutil.move(os.path.join(OUT, d, fn), os.path.join(dst, fn))
                shutil.rmtree(os.path.join(OUT, d))
    with open(os.path.join(OUT, "cases.json"), "w") as f:
        json.dump(CASES, f, indent=1)
    print(open(os.path.join(OUT, "B.db")).read())


if __name__ == "__main__":
    main()
