"""CPU check of the edit script of Compute_Trace_PTS (damar_b200/csrc/trace_script_core.cuh, the code the CUDA kernel runs one
segment per thread): tests/trace_script_host.cpp applies it to every record of committed reference .las files and the result
must equal what the UNMODIFIED reference function wrote (tests/golden/trace_script.json, made by oracle/_ref/trace_pts_ref
-- LAshow's call sequence around dalign/align.c:5577-5692 -- see tests/golden/make_trace_script.py)."""
import ctypes as C
import json
import os
import subprocess

import numpy as np
import pytest

from cases import GOLDEN, load_block
from damar_b200 import dbio

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SO = os.path.join(HERE, "_build", "libtracescript_host.so")
SRC = [os.path.join(HERE, "trace_script_host.cpp"), os.path.join(ROOT, "damar_b200", "csrc", "trace_script_core.cuh")]
with open(os.path.join(GOLDEN, "trace_script.json")) as _f:
    GOLD = json.load(_f)
with open(os.path.join(GOLDEN, "trace_mid.json")) as _f:        # Compute_Trace_MID (align.c:5694-5830) on the same files
    GOLD_MID = json.load(_f)


def _lib():
    if not os.path.exists(SO) or any(os.path.getmtime(SO) < os.path.getmtime(s) for s in SRC):
        os.makedirs(os.path.dirname(SO), exist_ok=True)
        subprocess.run(["g++", "-O2", "-g", "-std=c++17", "-fPIC", "-shared", "-Wall", "-o", SO, SRC[0]], check=True)
    L = C.CDLL(SO)
    for f in (L.ts_record_script, L.ts_record_mid):
        f.restype = C.c_int
        f.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int,
                      C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_int)]
    return L


def fnv(vals):
    h = 0xcbf29ce484222325
    for v in vals:
        h ^= int(v) & 0xffffffff
        h = (h * 0x100000001b3) & 0xffffffffffffffff
    return h


def complement_read(seq):
    return (3 - seq[::-1]).astype(np.int8)


def record_scripts(L, blk_of, las, fn="ts_record_script"):
    """yields (script, diffs) per record of a .las image; blk_of(read id) -> (block, local index)"""
    call = getattr(L, fn)
    novl, tspace, recs, traces = dbio.parse_las(las)
    tb = np.uint8 if tspace <= 125 else np.uint16
    out = np.zeros(1 << 20, np.int32)
    for r, t in zip(recs, traces):
        ab_, ai = blk_of(int(r["aread"]))
        bb_, bi = blk_of(int(r["bread"]))

        def seq(b, i, comp):
            s = b.bases[1 + int(b.boff[i]): 1 + int(b.boff[i]) + int(b.rlen[i])].astype(np.int8)
            if comp:
                s = complement_read(s)
            return np.concatenate([[4], s, [4]]).astype(np.int8)
        sa, sb = seq(ab_, ai, False), seq(bb_, bi, bool(r["flags"] & 1))
        tr = np.frombuffer(t, dtype=tb).astype(np.uint16)
        d = C.c_int(0)
        n = call(sa.ctypes.data + 1, len(sa) - 2, sb.ctypes.data + 1, len(sb) - 2, int(r["abpos"]), int(r["bbpos"]),
                 int(r["aepos"]), int(r["bepos"]), tr.ctypes.data, len(tr), tspace, out.ctypes.data, len(out), C.byref(d))
        assert n >= 0, n
        yield out[:n].copy(), d.value


@pytest.mark.parametrize("flavour", ["pts", "mid"])
@pytest.mark.parametrize("key", sorted(GOLD))
def test_edit_scripts_equal_reference(key, flavour):
    db, case, fn = key.split("/")
    root = {"tiny": "T", "rep": "R", "long": "L", "tan": "N"}[db]
    d = os.path.join(GOLDEN, db)
    _, nblocks, _, _ = dbio.read_stub(os.path.join(d, root + ".db"))
    blocks = [load_block(d, f"{root}.{i}", ()) for i in range(1, max(nblocks, 1) + 1)] if nblocks else [load_block(d, root, ())]

    def blk_of(rid):
        for b in blocks:
            if b.ufirst <= rid < b.ufirst + b.nreads:
                return b, rid - b.ufirst
        raise KeyError(rid)
    las = open(os.path.join(d, "ref", case, fn), "rb").read()
    g = (GOLD if flavour == "pts" else GOLD_MID)[key]
    L = _lib()
    n = 0
    for i, (scr, diffs) in enumerate(record_scripts(L, blk_of, las, "ts_record_script" if flavour == "pts" else "ts_record_mid")):
        tl, h, gd = g["records"][i]
        assert len(scr) == tl and "%016x" % fnv(scr) == h and diffs == gd, f"{key} record {i}: {len(scr)} vs {tl}"
        if str(i) in g["full"]:
            assert list(map(int, scr)) == g["full"][str(i)]
        n += 1
    assert n == g["n"]
