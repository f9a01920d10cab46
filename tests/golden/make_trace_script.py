#!/usr/bin/env python
"""Golden vectors for the edit script of Compute_Trace_PTS (dalign/align.c:5577-5692, iter_np :4892-5261, mode GREEDIEST as
utils/LAshow.c:261 calls it): for every record of committed reference .las files the length of the script the UNMODIFIED
reference function writes and a 64-bit FNV-1a hash of its values; every 400th record's script in full.

The same for Compute_Trace_MID (align.c:5694-5830, GREEDIEST as corrector/LAcorrect.c:545 calls it).

    python tests/golden/make_trace_script.py      ->  tests/golden/trace_script.json, tests/golden/trace_mid.json"""
import json
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
EXE = os.path.join(ROOT, "oracle", "_ref", "trace_pts_ref")
FILES = [("tiny", "T", "default", "T.1.T.2.las"), ("tiny", "T", "k16s200", "T.1.T.1.las"), ("rep", "R", "rep_default", "R.1.R.1.las"),
         ("rep", "R", "rep_s200", "R.1.R.1.las"), ("long", "L", "long_default", "L.1.L.1.las"), ("tan", "N", "tan_default", "N.1.N.1.las")]


def fnv(vals):
    h = 0xcbf29ce484222325
    for v in vals:
        h ^= v & 0xffffffff
        h = (h * 0x100000001b3) & 0xffffffffffffffff
    return h


def main():
    make("script", "trace_script.json")
    make("mid", "trace_mid.json")           # Compute_Trace_MID (align.c:5694-5830), same layout


def make(mode, target):
    out = {}
    for db, root, case, fn in FILES:
        las = os.path.join(HERE, db, "ref", case, fn)
        txt = subprocess.run([EXE, os.path.join(HERE, db, root), las, mode], check=True, capture_output=True, text=True).stdout
        recs, full = [], {}
        for i, ln in enumerate(txt.splitlines()):
            f = list(map(int, ln.split()))
            tl, scr = f[5], f[6:]
            assert len(scr) == tl
            recs.append([tl, "%016x" % fnv(scr), f[4]])
            if i % 400 == 0:
                full[str(i)] = scr
        out[f"{db}/{case}/{fn}"] = {"n": len(recs), "records": recs, "full": full}
        print(db, case, fn, len(recs), sum(r[0] for r in recs))
    with open(os.path.join(HERE, target), "w") as fh:
        json.dump(out, fh, separators=(",", ":"))


if __name__ == "__main__":
    main()
