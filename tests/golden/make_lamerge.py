#!/usr/bin/env python
"""Golden checksums for LAmerge (utils/LAmerge.c): the UNMODIFIED reference LAmerge (oracle/_ref/LAmerge) merges committed
reference .las files; its output is recorded as size + sha256 (the merge copies records verbatim, padding bytes included,
so a byte-exact merge of the same inputs has the same digest).

    python tests/golden/make_lamerge.py      ->  tests/golden/lamerge.json"""
import hashlib
import json
import os
import shutil
import subprocess
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
EXE = os.path.join(ROOT, "oracle", "_ref", "LAmerge")
CASES = {
    "T.1": ("tiny", "T", ["default/T.1.T.1.las", "default/T.1.T.2.las"]),
    "T.2": ("tiny", "T", ["b2/T.2.T.1.las", "b2/T.2.T.2.las"]),
    "T.2.rev": ("tiny", "T", ["b2/T.2.T.2.las", "b2/T.2.T.1.las"]),                       # input order must not matter
    "T.1.s200": ("tiny", "T", ["k16s200/T.1.T.1.las"]),                                    # one file, 16-bit traces
    "R.1": ("rep", "R", ["rep_default/R.1.R.1.las", "rep_default/R.1.R.2.las"]),
    "E.F": ("edge", "F", ["edge_cross/F.1.E.1.las", "edge_cross/F.1.F.1.las", "edge_none/F.1.F.1.las"]),   # empty files
}


def main():
    out = {}
    for name, (db, root, files) in CASES.items():
        with tempfile.TemporaryDirectory() as tmp:
            for fn in os.listdir(os.path.join(HERE, db)):
                p = os.path.join(HERE, db, fn)
                if os.path.isfile(p):
                    shutil.copy(p, tmp)
            ins = []
            for i, f in enumerate(files):
                dst = os.path.join(tmp, f"in{i}.las")
                shutil.copy(os.path.join(HERE, db, "ref", f), dst)
                ins.append(dst)
            r = subprocess.run([EXE, root, "out.las"] + [os.path.basename(x) for x in ins], cwd=tmp, capture_output=True, text=True)
            assert r.returncode == 0, r.stderr
            data = open(os.path.join(tmp, "out.las"), "rb").read()
        out[name] = {"db": db, "inputs": files, "size": len(data), "sha256": hashlib.sha256(data).hexdigest()}
        print(name, len(data))
    with open(os.path.join(HERE, "lamerge.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
