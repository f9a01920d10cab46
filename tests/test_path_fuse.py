"""CPU check of Handle_Redundancies (damar_b200/csrc/path_fuse_core.cuh, the code the CUDA kernel runs one pair per thread): the oracle buffers the UNFUSED
alignments of every pair (it restates the path up to filter.c:2436), tests/path_fuse_host.cpp applies the
product's post-pass to them, and the serialised result must equal the reference daligner's .las files of the
`rep` fixture (tests/golden/make_repeats.py: containments, fusions and bridges)."""
import ctypes as C
import json
import os
import subprocess

import pytest

import oracle_ffi as O
from cases import GOLDEN, load_block, oracle_kwargs, parse_flags
from damar_b200 import dbio

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SO = os.path.join(HERE, "_build", "libpathfuse_host.so")
SRC = [os.path.join(HERE, "path_fuse_host.cpp"), os.path.join(ROOT, "damar_b200", "csrc", "path_fuse_core.cuh")]
REP = os.path.join(GOLDEN, "rep")
with open(os.path.join(REP, "cases.json")) as _f:
    REP_CASES = json.load(_f)


def _lib():
    if not os.path.exists(SO) or any(os.path.getmtime(SO) < os.path.getmtime(s) for s in SRC):
        os.makedirs(os.path.dirname(SO), exist_ok=True)
        subprocess.run(["g++", "-O2", "-g", "-std=c++17", "-fPIC", "-shared", "-Wall", "-x", "c++", "-o", SO, SRC[0]], check=True)
    L = C.CDLL(SO)
    L.pf_fuse_result.restype = C.c_int64
    L.pf_fuse_result.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
    L.pf_fuse_result_self.restype = C.c_int64
    L.pf_fuse_result_self.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int]
    return L


def oracle_daligner_fused(ablk, bblks, kw, fuse):
    """tests/oracle_ffi.py::daligner with the post-pass applied after every ora_match_filter call."""
    L = O.lib()
    P = _lib()
    out, changed = {}, 0
    ha = O.HostBlock(ablk)
    kw = dict(kw)
    kw["freq"] = tuple(float(x) for x in ablk.freq)
    p = O.make_params(**kw)
    aptr, alen, _ = O.sort_kmers(ha, p)
    for bblk in bblks:
        res = L.ora_new_result()
        selfcmp = bblk is ablk
        hb = ha if selfcmp else O.HostBlock(bblk)
        p.sizeof_a, p.sizeof_b = ablk.sizeof_db(), bblk.sizeof_db()
        bptr, blen = (aptr, alen) if selfcmp else O.sort_kmers(hb, p)[:2]
        L.ora_match_filter(C.byref(ha.c), C.byref(hb.c), aptr, alen, bptr, blen, int(selfcmp), 0, C.byref(p), res)
        if fuse:
            changed += P.pf_fuse_result(res, 0, C.byref(ha.c), C.byref(hb.c), int(selfcmp), 0, p.symmetric, p.tspace)
        mark = res.contents.novl
        if not selfcmp and bptr:
            L.ora_free(bptr)
        cb = L.ora_complement_block(C.byref(hb.c))
        n = C.c_int64(0)
        cptr = L.ora_sort_kmers(cb, C.byref(p), C.byref(n))
        L.ora_match_filter(C.byref(ha.c), cb, aptr, alen, cptr, n.value, int(selfcmp), 1, C.byref(p), res)
        if fuse:
            changed += P.pf_fuse_result(res, mark, C.byref(ha.c), cb, int(selfcmp), 1, p.symmetric, p.tspace)
        if cptr:
            L.ora_free(cptr)
        L.ora_free_block(cb)
        an, bn = ablk.name, bblk.name
        if selfcmp or not p.symmetric:
            out[f"{an}.{bn}.las"] = O.serialise(res, p, 0, -1)[0]
        else:
            last = (bblk.ufirst + bblk.nreads - 1) if bblk.part < ablk.part else (ablk.ufirst + ablk.nreads - 1)
            b1, b2 = O.serialise(res, p, 0, last)
            n1, n2 = f"{an}.{bn}.las", f"{bn}.{an}.las"
            if bblk.part < ablk.part:
                n1, n2 = n2, n1
            out[n1], out[n2] = b1, b2
        L.ora_free_result(res)
    if aptr:
        L.ora_free(aptr)
    return out, changed


@pytest.mark.parametrize("case", sorted(REP_CASES))
def test_fused_oracle_output_equals_reference(case):
    o = parse_flags(REP_CASES[case]["flags"])
    names = REP_CASES[case]["blocks"]
    blks = {n: load_block(REP, n, ()) for n in set(names)}
    ablk, bblks = blks[names[0]], [blks[n] for n in names[1:]]
    got, changed = oracle_daligner_fused(ablk, bblks, oracle_kwargs(o), fuse=True)
    assert changed > 0                                       # the fixture does exercise the post-pass
    d = os.path.join(REP, "ref", case)
    for fn in sorted(os.listdir(d)):
        ref = open(os.path.join(d, fn), "rb").read()
        ok, why = dbio.las_equal(ref, got[fn])
        assert ok, f"{case}/{fn}: {why}"


def test_unfused_output_differs():
    """Sanity: without the post-pass the same fixture does NOT reproduce the reference (so the test above is not vacuous)."""
    case = "rep_default"
    o = parse_flags(REP_CASES[case]["flags"])
    names = REP_CASES[case]["blocks"]
    blks = {n: load_block(REP, n, ()) for n in set(names)}
    got, _ = oracle_daligner_fused(blks[names[0]], [blks[n] for n in names[1:]], oracle_kwargs(o), fuse=False)
    d = os.path.join(REP, "ref", case)
    assert any(not dbio.las_equal(open(os.path.join(d, fn), "rb").read(), got[fn])[0] for fn in os.listdir(d))


from cases import TAN, TAN_CASES  # noqa: E402


@pytest.mark.parametrize("case", sorted(TAN_CASES))
def test_oracle_datander_equals_reference(case):
    """datander (scrub/tandem.c): the oracle's Match_Self restatement + the product's Handle_Redundancies pass without
    bridging (tandem.c:767-850) == the reference datander's .las on reads with tandem arrays (tests/golden/make_tandem.py)."""
    o = O.parse_tandem_flags(TAN_CASES[case]["flags"])
    blk = load_block(TAN, TAN_CASES[case]["blocks"][0], ())
    got, changed = O.datander(blk, o, fuse=_lib().pf_fuse_result_self)
    d = os.path.join(TAN, "ref", case)
    for fn in sorted(os.listdir(d)):
        ref = open(os.path.join(d, fn), "rb").read()
        ok, why = dbio.las_equal(ref, got[fn])
        assert ok, f"{case}/{fn}: {why}"
    if case == "tan_b2":
        assert changed > 0                                   # block 2 has reads whose alignments fuse / contain each other
