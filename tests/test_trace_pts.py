"""Compute_Trace_PTS (dalign/align.c:5577-5692) in its cost form.

CPU : the oracle's textbook dynamic program (oracle/daligner_oracle.c::ora_trace_point_diffs) reproduces path->diffs as the
      UNMODIFIED reference function leaves it (tests/golden/trace_pts.json, made by tests/golden/make_trace_pts.py), including
      the bridged records of the `rep` fixture whose recorded diffs are not optimal.
GPU : the CUDA kernel (damar_trace_point_diffs through the C-ABI) equals the golden values on every record and the oracle on
      every segment, and shows that every (d_i, b_i) the wave kernel emitted is realisable: exact distance <= d_i."""
import json
import os

import numpy as np
import pytest

import oracle_ffi as O
from cases import GOLDEN, load_block
from damar_b200 import dbio

with open(os.path.join(GOLDEN, "trace_pts.json")) as _f:
    GOLD = json.load(_f)


def _blocks(key):
    db, case, fn = key.split("/")
    a, b = fn.split(".")[1], fn.split(".")[3]
    root = fn.split(".")[0]
    d = os.path.join(GOLDEN, db)
    return load_block(d, f"{root}.{a}", ()), load_block(d, f"{root}.{b}", ()), open(os.path.join(d, "ref", case, fn), "rb").read()


@pytest.mark.parametrize("key", sorted(GOLD))
def test_oracle_matches_reference_compute_trace_pts(key):
    ablk, bblk, las = _blocks(key)
    want = np.array(GOLD[key]["diffs"])
    novl, tspace, recs, _ = dbio.parse_las(las)
    assert novl == GOLD[key]["n"]
    changed = np.nonzero(want != recs["diffs"])[0]
    assert len(changed) == GOLD[key]["changed"]
    every = 6
    idx, rec, seg, soff = O.trace_point_diffs(ablk, bblk, las, every=every)
    assert np.array_equal(rec, want[idx])
    if len(changed):                                       # the records where re-alignment finds fewer differences
        sub = bytearray(las[:12])
        import struct
        parts, o = [], 12
        for i in range(novl):
            n = 40 + int(recs[i]["tlen"]) * (1 if tspace <= 125 else 2)
            if i in set(changed.tolist()):
                parts.append(las[o:o + n])
            o += n
        sub[:8] = struct.pack("<q", len(parts))
        _, rec2, _, _ = O.trace_point_diffs(ablk, bblk, bytes(sub) + b"".join(parts))
        assert np.array_equal(rec2, want[changed])


@pytest.mark.gpu
@pytest.mark.parametrize("key", sorted(GOLD))
def test_cuda_trace_point_diffs(key):
    from damar_b200 import api
    api.init()
    ablk, bblk, las = _blocks(key)
    want = np.array(GOLD[key]["diffs"])
    da = api.DeviceBlockHandle(ablk)
    db = da if bblk.name == ablk.name else api.DeviceBlockHandle(bblk)
    rec, seg, soff = api.trace_point_diffs(da, db, las, segments=True)
    assert np.array_equal(rec, want)                       # == the reference's Compute_Trace_PTS, every record
    novl, tspace, recs, traces = dbio.parse_las(las)
    # every emitted trace-point segment is realisable with at most its recorded number of differences
    tb = np.uint8 if tspace <= 125 else np.uint16
    d_rec = np.concatenate([np.frombuffer(t, dtype=tb)[0::2].astype(np.int64) for t in traces])
    assert len(d_rec) == len(seg) and np.all(seg <= d_rec)
    assert int((seg < d_rec).sum()) >= (1 if GOLD[key]["changed"] else 0)
    # per segment against the oracle on a sample of the records
    idx, orec, oseg, ooff = O.trace_point_diffs(ablk, bblk, las, every=9)
    for j, i in enumerate(idx):
        n = int(recs[i]["tlen"]) // 2
        assert np.array_equal(seg[soff[i]: soff[i] + n], oseg[ooff[j]: ooff[j] + n]), f"record {i}"
    if db is not da:
        db.free()
    da.free()


@pytest.mark.gpu
def test_cuda_trace_point_diffs_rejects_out_of_bounds():
    """TP_Error (align.c:5667): a record whose end lies beyond the read is refused, the others are still answered."""
    from damar_b200 import api
    api.init()
    key = "tiny/default/T.1.T.1.las"
    ablk, bblk, las = _blocks(key)
    novl, tspace, recs, traces = dbio.parse_las(las)
    buf = bytearray(las)
    import struct
    struct.pack_into("<i", buf, 12 + 16, 10 ** 7)          # aepos of record 0
    da = api.DeviceBlockHandle(ablk)
    rec = api.trace_point_diffs(da, da, bytes(buf))
    assert rec[0] == -1 and np.array_equal(rec[1:], np.array(GOLD[key]["diffs"])[1:])
    da.free()


@pytest.mark.gpu
@pytest.mark.parametrize("middles", [False, True])
@pytest.mark.parametrize("key", ["tiny/default/T.1.T.2.las", "rep/rep_default/R.1.R.1.las", "long/long_default/L.1.L.1.las", "tan/tan_default/N.1.N.1.las"])
def test_gpu_edit_scripts_equal_reference(key, middles):
    """damar_trace_point_scripts / damar_trace_mid_scripts (trace_pts.cu + trace_script_core.cuh) against the output of the
    UNMODIFIED Compute_Trace_PTS / Compute_Trace_MID (tests/golden/trace_script.json, trace_mid.json): script length and hash
    of every record, every 400th script in full, and the diffs."""
    import json
    from cases import GOLDEN, load_block
    from damar_b200 import api, dbio
    from test_trace_script import fnv
    with open(os.path.join(GOLDEN, "trace_mid.json" if middles else "trace_script.json")) as f:
        g = json.load(f)[key]
    db, case, fn = key.split("/")
    root = {"tiny": "T", "rep": "R", "long": "L", "tan": "N"}[db]
    an, bn = ".".join(fn.split(".")[0:2]), ".".join(fn.split(".")[2:4])
    api.init()
    da = api.DeviceBlockHandle(load_block(os.path.join(GOLDEN, db), an, ()))
    dbh = da if bn == an else api.DeviceBlockHandle(load_block(os.path.join(GOLDEN, db), bn, ()))
    las = open(os.path.join(GOLDEN, db, "ref", case, fn), "rb").read()
    rec, soff, script = api.trace_point_scripts(da, dbh, las, middles=middles)
    assert len(rec) == g["n"] and soff[-1] == len(script)
    for i, (tl, h, gd) in enumerate(g["records"]):
        scr = script[soff[i]: soff[i + 1]]
        assert len(scr) == tl and rec[i] == gd, (i, len(scr), tl, rec[i], gd)
        assert "%016x" % fnv(scr) == h, i
        if str(i) in g["full"]:
            assert list(map(int, scr)) == g["full"][str(i)]
    da.free()
    if dbh is not da:
        dbh.free()
