#!/usr/bin/env python
"""Golden vectors for Compute_Trace_PTS (dalign/align.c:5577-5692): path->diffs of every record of committed reference
.las files after the UNMODIFIED reference function ran on it (oracle/_ref/trace_pts_ref = oracle/ref_trace_pts.c, a driver
with LAshow's call sequence, linked with the reference's align.c and DB.c).

    python tests/golden/make_trace_pts.py      ->  tests/golden/trace_pts.json

The `rep` files hold fused / bridged records (Handle_Redundancies), whose recorded diffs are NOT all optimal: there the
reference's values differ from the records' own, which is what makes the vectors discriminating."""
import json
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
EXE = os.path.join(ROOT, "oracle", "_ref", "trace_pts_ref")
FILES = [("tiny", "T", "default", "T.1.T.1.las"), ("tiny", "T", "default", "T.1.T.2.las"), ("tiny", "T", "default", "T.2.T.1.las"),
         ("tiny", "T", "k16s200", None), ("rep", "R", "rep_default", "R.1.R.1.las"), ("rep", "R", "rep_default", "R.1.R.2.las"),
         ("rep", "R", "rep_s200", "R.1.R.1.las")]


def main():
    out = {}
    for db, root, case, fn in FILES:
        d = os.path.join(HERE, db, "ref", case)
        for f in ([fn] if fn else sorted(os.listdir(d))):
            txt = subprocess.run([EXE, os.path.join(HERE, db, root), os.path.join(d, f)], check=True, capture_output=True, text=True).stdout
            rows = [tuple(map(int, ln.split())) for ln in txt.splitlines()]
            out[f"{db}/{case}/{f}"] = {"n": len(rows), "changed": sum(1 for r in rows if r[3] != r[4]), "diffs": [r[4] for r in rows]}
            print(db, case, f, len(rows), out[f"{db}/{case}/{f}"]["changed"])
    with open(os.path.join(HERE, "trace_pts.json"), "w") as fh:
        json.dump(out, fh, separators=(",", ":"))


if __name__ == "__main__":
    main()
