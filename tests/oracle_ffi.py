"""ctypes binding of oracle/libdaligner_oracle.so -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
REF_DIR = os.path.join(ORACLE_DIR, "_ref")


class OraBlock(C.Structure):
    _fields_ = [("nreads", C.c_int32), ("maxlen", C.c_int32), ("totlen", C.c_int64),
                ("ufirst", C.c_int32), ("bases", C.c_void_p), ("boff", C.c_void_p),
                ("rlen", C.c_void_p), ("manno", C.c_void_p), ("mdata", C.c_void_p)]


class OraParams(C.Structure):
    _fields_ = [("kmer", C.c_int32), ("binshift", C.c_int32), ("suppress", C.c_int32),
                ("hitmin", C.c_int32), ("nthreads", C.c_int32), ("identity", C.c_int32),
                ("symmetric", C.c_int32), ("minover", C.c_int32), ("tspace", C.c_int32),
                ("no_trace", C.c_int32), ("ave_corr", C.c_double), ("freq", C.c_float * 4),
                ("mem_limit", C.c_uint64), ("sizeof_a", C.c_int64), ("sizeof_b", C.c_int64),
                ("biased", C.c_int32)]


class OraOvl(C.Structure):
    _fields_ = [("tlen", C.c_int32), ("diffs", C.c_int32), ("abpos", C.c_int32), ("bbpos", C.c_int32),
                ("aepos", C.c_int32), ("bepos", C.c_int32), ("flags", C.c_uint32), ("aread", C.c_int32),
                ("bread", C.c_int32), ("toff", C.c_int64)]


class OraResult(C.Structure):
    _fields_ = [("novl", C.c_int64), ("omax", C.c_int64), ("ovl", C.POINTER(OraOvl)),
                ("ttop", C.c_int64), ("tmax", C.c_int64), ("trace", C.POINTER(C.c_uint16)),
                ("nhits", C.c_int64), ("nfilt", C.c_int64), ("ncheck", C.c_int64),
                ("multi_la_pairs", C.c_int64), ("wave_cells", C.c_int64), ("waves", C.c_int64),
                ("limit", C.c_int32)]


KMER_DT = np.dtype([("code", "<u8"), ("rpos", "<i4"), ("read", "<i4")])
HIT_DT = np.dtype([("diag", "<i4"), ("apos", "<i4"), ("aread", "<i4"), ("bread", "<i4")])

_lib = None


def build() -> None:
    subprocess.run(["make", "-s", "-C", ORACLE_DIR, "port"], check=True)


def lib():
    global _lib
    if _lib is None:
        so = os.path.join(ORACLE_DIR, "libdaligner_oracle.so")
        src = os.path.join(ORACLE_DIR, "daligner_oracle.c")
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
            build()
        L = C.CDLL(so)
        L.ora_sort_kmers.restype = C.c_void_p
        L.ora_sort_kmers.argtypes = [C.POINTER(OraBlock), C.POINTER(OraParams), C.POINTER(C.c_int64)]
        L.ora_seed_hits.restype = C.c_void_p
        L.ora_seed_hits.argtypes = [C.POINTER(OraBlock), C.POINTER(OraBlock), C.c_void_p, C.c_int64,
                                    C.c_void_p, C.c_int64, C.c_int, C.c_int, C.POINTER(OraParams),
                                    C.POINTER(C.c_int64), C.POINTER(C.c_int32)]
        L.ora_report.restype = None
        L.ora_report.argtypes = [C.POINTER(OraBlock), C.POINTER(OraBlock), C.c_void_p, C.c_int64,
                                 C.c_int, C.c_int, C.POINTER(OraParams), C.POINTER(OraResult)]
        L.ora_match_filter.restype = None
        L.ora_match_filter.argtypes = [C.POINTER(OraBlock), C.POINTER(OraBlock), C.c_void_p, C.c_int64,
                                       C.c_void_p, C.c_int64, C.c_int, C.c_int, C.POINTER(OraParams),
                                       C.POINTER(OraResult)]
        L.ora_match_self.restype = None
        L.ora_match_self.argtypes = [C.POINTER(OraBlock), C.POINTER(OraParams), C.POINTER(OraResult)]
        L.ora_complement_block.restype = C.POINTER(OraBlock)
        L.ora_complement_block.argtypes = [C.POINTER(OraBlock)]
        L.ora_free_block.argtypes = [C.POINTER(OraBlock)]
        L.ora_new_result.restype = C.POINTER(OraResult)
        L.ora_free_result.argtypes = [C.POINTER(OraResult)]
        L.ora_reset_result.argtypes = [C.POINTER(OraResult)]
        L.ora_serialise_las.restype = None
        L.ora_serialise_las.argtypes = [C.POINTER(OraResult), C.POINTER(OraParams), C.c_int, C.c_int,
                                        C.POINTER(C.c_void_p), C.POINTER(C.c_int64),
                                        C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]
        L.ora_local_alignment.restype = None
        L.ora_local_alignment.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                          C.c_int, C.c_int, C.POINTER(OraParams), C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_void_p]
        L.ora_trace_point_diffs.restype = None
        L.ora_trace_point_diffs.argtypes = [C.POINTER(OraBlock), C.POINTER(OraBlock), C.POINTER(OraBlock), C.c_void_p, C.c_void_p,
                                            C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p]
        L.ora_free.argtypes = [C.c_void_p]
        _lib = L
    return _lib


class HostBlock:
    """Keeps the numpy arrays alive behind an OraBlock."""

    def __init__(self, blk):
        self.blk = blk
        self.bases = np.ascontiguousarray(blk.bases, dtype=np.uint8)
        self.boff = np.ascontiguousarray(blk.boff, dtype=np.int64)
        self.rlen = np.ascontiguousarray(blk.rlen, dtype=np.int32)
        self.manno = None if blk.mask_anno is None else np.ascontiguousarray(blk.mask_anno, dtype=np.int64)
        self.mdata = None if blk.mask_data is None else np.ascontiguousarray(
            blk.mask_data if blk.mask_data.size else np.zeros(1, np.int32), dtype=np.int32)
        self.c = OraBlock(blk.nreads, blk.maxlen, blk.totlen, blk.ufirst,
                          self.bases.ctypes.data + 1, self.boff.ctypes.data, self.rlen.ctypes.data,
                          None if self.manno is None else self.manno.ctypes.data,
                          None if self.mdata is None else self.mdata.ctypes.data)


def make_params(k=14, w=6, h=35, t=0, nthreads=4, identity=0, symmetric=1, minlen=1000, tspace=100,
                no_trace=0, ave_corr=0.70, freq=(.25, .25, .25, .25), mem_limit=0, sizeof_a=0, sizeof_b=0, biased=0):
    p = OraParams()
    p.kmer, p.binshift, p.suppress, p.hitmin, p.nthreads = k, w, t, h, nthreads
    p.identity, p.symmetric, p.minover, p.tspace, p.no_trace = identity, symmetric, 2 * minlen, tspace, no_trace
    p.ave_corr = ave_corr
    for i in range(4):
        p.freq[i] = float(freq[i])
    p.mem_limit, p.sizeof_a, p.sizeof_b = mem_limit, sizeof_a, sizeof_b
    p.biased = biased
    return p


def sort_kmers(hb: HostBlock, p: OraParams):
    """Returns (ptr, len, numpy view incl. the two sentinels)."""
    n = C.c_int64(0)
    ptr = lib().ora_sort_kmers(C.byref(hb.c), C.byref(p), C.byref(n))
    if not ptr:
        return None, 0, np.zeros(0, KMER_DT)
    arr = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=((n.value + 2) * 16,)).view(KMER_DT)
    return ptr, n.value, arr


def seed_hits(ha: HostBlock, hb: HostBlock, aptr, alen, bptr, blen, self_, comp, p):
    n = C.c_int64(0)
    lim = C.c_int32(0)
    ptr = lib().ora_seed_hits(C.byref(ha.c), C.byref(hb.c), aptr, alen, bptr, blen, self_, comp,
                              C.byref(p), C.byref(n), C.byref(lim))
    if not ptr:
        return None, 0, np.zeros(0, HIT_DT), lim.value
    arr = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=((n.value + 1) * 16,)).view(HIT_DT)
    return ptr, n.value, arr, lim.value


def serialise(res, p, only_identity=0, split_read=-1):
    o1, o2 = C.c_void_p(), C.c_void_p()
    l1, l2 = C.c_int64(0), C.c_int64(0)
    lib().ora_serialise_las(res, C.byref(p), only_identity, split_read, C.byref(o1), C.byref(l1),
                            C.byref(o2), C.byref(l2))
    b1 = C.string_at(o1, l1.value)
    lib().ora_free(o1)
    b2 = None
    if o2.value:
        b2 = C.string_at(o2, l2.value)
        lib().ora_free(o2)
    return b1, b2


def daligner(ablk, bblks, p_kwargs, only_identity=0):
    """Oracle equivalent of `daligner [opts] A B1 B2 ...` (daligner.c:958-1073).
    Returns {filename: bytes} keyed like the reference's output files
    ('<root>.<a>.<root>.<b>.las')."""
    L = lib()
    out = {}
    ha = HostBlock(ablk)
    kw = dict(p_kwargs)
    kw["freq"] = tuple(float(x) for x in ablk.freq)
    p = make_params(**kw)
    aptr, alen, _ = sort_kmers(ha, p)
    for bblk in bblks:
        res = L.ora_new_result()
        selfcmp = bblk is ablk
        hb = ha if selfcmp else HostBlock(bblk)
        p.sizeof_a, p.sizeof_b = ablk.sizeof_db(), bblk.sizeof_db()
        if selfcmp:
            bptr, blen = aptr, alen
        else:
            bptr, blen, _ = sort_kmers(hb, p)
        L.ora_match_filter(C.byref(ha.c), C.byref(hb.c), aptr, alen, bptr, blen, int(selfcmp), 0, C.byref(p), res)
        if not selfcmp and bptr:
            L.ora_free(bptr)
        cb = L.ora_complement_block(C.byref(hb.c))
        n = C.c_int64(0)
        cptr = L.ora_sort_kmers(cb, C.byref(p), C.byref(n))
        L.ora_match_filter(C.byref(ha.c), cb, aptr, alen, cptr, n.value, int(selfcmp), 1, C.byref(p), res)
        if cptr:
            L.ora_free(cptr)
        L.ora_free_block(cb)
        stats = dict(nfilt=res.contents.nfilt, ncheck=res.contents.ncheck, multi=res.contents.multi_la_pairs,
                     cells=res.contents.wave_cells, waves=res.contents.waves, novl=res.contents.novl)
        an, bn = ablk.name, bblk.name
        if selfcmp or not p.symmetric:
            b1, _ = serialise(res, p, only_identity, -1)
            out[f"{an}.{bn}.las"] = b1
        else:
            if bblk.part < ablk.part:
                last = bblk.ufirst + bblk.nreads - 1
            else:
                last = ablk.ufirst + ablk.nreads - 1
            b1, b2 = serialise(res, p, only_identity, last)
            n1, n2 = f"{an}.{bn}.las", f"{bn}.{an}.las"
            if bblk.part < ablk.part:
                n1, n2 = n2, n1
            out[n1], out[n2] = b1, b2
        out.setdefault("_stats", []).append(stats)
        L.ora_free_result(res)
    if aptr:
        L.ora_free(aptr)
    return out


def parse_tandem_flags(flags):
    """datander's options (scrub/datander.c:142-176): -k12 -w4 -h35 -e.70 -l500 -s100 -j4."""
    o = dict(k=12, w=4, h=35, e=0.70, l=500, s=100, j=4)
    for f in flags:
        c, v = f[1], f[2:]
        if c in "kwhlsj":
            o[c] = int(v)
        elif c == "e":
            o["e"] = float(v)
        elif c != "v":
            raise ValueError(f)
    return o


def datander(blk, o, fuse=None):
    """Oracle equivalent of `datander [opts] <block>` (scrub/datander.c:236-256): {file name: bytes}.  `fuse` = the
    pf_fuse_result_self entry of tests/path_fuse_host.cpp (Handle_Redundancies is the product's code run on the host)."""
    L = lib()
    hb = HostBlock(blk)
    p = make_params(k=o["k"], w=o["w"], h=o["h"], nthreads=o["j"], minlen=o["l"], tspace=o["s"], ave_corr=o["e"],
                    freq=tuple(float(x) for x in blk.freq))
    res = L.ora_new_result()
    L.ora_match_self(C.byref(hb.c), C.byref(p), res)
    changed = fuse(res, 0, C.byref(hb.c), p.tspace) if fuse is not None else 0
    b1, _ = serialise(res, p, 0, -1)
    L.ora_free_result(res)
    return {f"{blk.name}.{blk.name}.las": b1}, changed


def trace_point_diffs(ablk, bblk, las: bytes, every: int = 1):
    """ora_trace_point_diffs over (every `every`-th record of) a .las image: (record indices, per-record sums,
    per-segment values, first-segment offsets into the per-segment array of the SELECTED records)."""
    from damar_b200 import dbio
    novl, tspace, recs, traces = dbio.parse_las(las)
    idx = np.arange(0, novl, every)
    tbytes = 1 if tspace <= 125 else 2
    sel = np.ascontiguousarray(recs[idx])
    tr = [np.frombuffer(traces[i], dtype=np.uint8 if tbytes == 1 else np.uint16).astype(np.uint16) for i in idx]
    toff = np.zeros(len(idx) + 1, np.uint64)
    toff[1:] = np.cumsum([len(t) for t in tr])
    trace = np.concatenate(tr) if tr else np.zeros(1, np.uint16)
    ha, hb = HostBlock(ablk), HostBlock(bblk)
    L = lib()
    cb = L.ora_complement_block(C.byref(hb.c))
    rec = np.zeros(max(len(idx), 1), np.int32)
    seg = np.zeros(max(int(toff[-1]) // 2, 1), np.uint16)
    L.ora_trace_point_diffs(C.byref(ha.c), C.byref(hb.c), cb, sel.ctypes.data, toff.ctypes.data, trace.ctypes.data,
                            len(idx), tspace, rec.ctypes.data, seg.ctypes.data)
    L.ora_free_block(cb)
    return idx, rec[: len(idx)], seg[: int(toff[-1]) // 2], (toff[:-1] // 2).astype(np.int64)
