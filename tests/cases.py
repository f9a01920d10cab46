"""Shared helpers: golden cases, flag parsing, DB loading (test infrastructure)."""
from __future__ import annotations

import json
import os
from functools import lru_cache

from damar_b200 import dbio

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
TINY = os.path.join(GOLDEN, "tiny")

with open(os.path.join(GOLDEN, "cases.json")) as f:
    CASES = json.load(f)


BIAS = os.path.join(GOLDEN, "bias")        # AT-rich DB for -b (tests/golden/make_biased.py)
with open(os.path.join(BIAS, "cases.json")) as f:
    BIAS_CASES = json.load(f)


EDGE = os.path.join(GOLDEN, "edge")        # corner cases (tests/golden/make_edge.py)
with open(os.path.join(EDGE, "cases.json")) as f:
    EDGE_CASES = json.load(f)


GRAM = os.path.join(GOLDEN, "gram")        # MAXGRAM cap, oversized runs, > 127 alignments per pair (tests/golden/make_gram.py)
with open(os.path.join(GRAM, "cases.json")) as f:
    GRAM_CASES = json.load(f)


LONG = os.path.join(GOLDEN, "long")        # reads of 34-86 kbp at 15 % error, -e.65 (tests/golden/make_long.py)
with open(os.path.join(LONG, "cases.json")) as f:
    LONG_CASES = json.load(f)


TAN = os.path.join(GOLDEN, "tan")          # datander: reads with tandem arrays (tests/golden/make_tandem.py)
with open(os.path.join(TAN, "cases.json")) as f:
    TAN_CASES = json.load(f)


def parse_flags(flags):
    """daligner flags -> dict(k,w,h,t,e,l,s,j,M,masks,I,O,A,T) (daligner.c:721-808)."""
    o = dict(k=14, w=6, h=35, t=0, e=0.70, l=1000, s=100, j=4, M=None, masks=[], I=False, O=False, A=False, T=False, b=False)
    for f in flags:
        c, v = f[1], f[2:]
        if c in "kwhtlsjM":
            o[c] = int(v)
        elif c == "e":
            o["e"] = float(v)
        elif c == "m":
            o["masks"].append(v)
        elif c in "IOATb":
            o[c] = True
        elif c == "v":
            pass
        else:
            raise ValueError(f)
    return o


def _phys():
    from damar_b200.api import physical_memory
    return physical_memory()


def oracle_kwargs(o):
    return dict(k=o["k"], w=o["w"], h=o["h"], t=o["t"], nthreads=o["j"], identity=int(o["I"] or o["O"]),
                symmetric=int(not o["A"]), minlen=o["l"], tspace=o["s"], no_trace=int(o["T"]), ave_corr=o["e"],
                mem_limit=_phys() if o["M"] is None else o["M"] << 30, biased=int(o.get("b", False)))


def settings(o):
    from damar_b200.api import Settings
    return Settings(kmer=o["k"], binshift=o["w"], hitmin=o["h"], suppress=o["t"], ave_corr=o["e"], minlen=o["l"],
                    trace_spacing=o["s"], nthreads=o["j"], mem_limit_gb=o["M"], symmetric=not o["A"],
                    identity=o["I"], only_identity=o["O"], no_trace=o["T"], biased=o.get("b", False))


@lru_cache(maxsize=None)
def load_block(dbdir, name, masks):
    return dbio.read_block(os.path.join(dbdir, name), list(masks) if masks else None)


def golden_files(case, dbdir=TINY):
    d = os.path.join(dbdir, "ref", case)
    return {fn: open(os.path.join(d, fn), "rb").read() for fn in sorted(os.listdir(d))}
