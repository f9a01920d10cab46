"""Dazzler-DB block, mask-track and .las I/O (host-side mirror of the reference loaders).

Formats follow the reference byte for byte (paths relative to the reference root):
  * ``.db`` stub / ``.idx`` / ``.bps``  -- db/DB.c:462-672 (Open_DB_Block), db/DB.c:1547-1608
    (Read_All_Sequences), db/DB.c:358-381 (Uncompress_Read); HITS_DB is 88 bytes on disk and
    HITS_READ 32 bytes (db/DB.h:319-325, 361-389).
  * Dazzler interval tracks ``.<db>.<track>.anno/.data`` -- db/DB.c:1113-1294 (Load_Track) and the
    ``anno[j] /= sizeof(int)`` step of daligner.c:480-482.
  * ``.las`` -- align.c:3345-3373 (Write_Overlap), lib.python/marvel/LAS.py:35-37.

This module is plain numpy; it holds no alignment arithmetic.
"""
from __future__ import annotations

import os
import re
import struct
from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import numpy as np

HITS_DB_BYTES = 88
HITS_READ_BYTES = 32
_READ_DT = np.dtype([("rlen", "<i4"), ("_p0", "<i4"), ("boff", "<i8"), ("coff", "<i8"),
                     ("flags", "<i4"), ("_p1", "<i4")])
assert _READ_DT.itemsize == HITS_READ_BYTES


@dataclass
class Block:
    """One DB block in core: the arrays Read_All_Sequences leaves behind."""
    name: str                 # e.g. "G.1" (what daligner prints / uses in file names)
    root: str                 # e.g. "G"
    path: str                 # "<dir>/.G"  (HITS_DB.path; prefix for .bps/.idx/tracks)
    part: int
    ufirst: int
    nreads: int
    maxlen: int
    totlen: int
    ureads: int
    freq: np.ndarray          # float32[4]
    rlen: np.ndarray          # int32[nreads]
    boff: np.ndarray          # int64[nreads+1] in-memory offsets (cumulative rlen+1)
    bases: np.ndarray         # uint8[totlen+nreads+1]; bases[0] is the leading 4, read i at 1+boff[i]
    mask_anno: Optional[np.ndarray] = None   # int64[nreads+1], units of int32
    mask_data: Optional[np.ndarray] = None   # int32[...], sorted [beg,end) pairs per read
    track_names: List[str] = field(default_factory=list)

    def sizeof_db(self) -> int:
        """sizeof_DB (db/DB.c:726-756) for the limit rule of filter.c:2645."""
        s = HITS_DB_BYTES + HITS_READ_BYTES * (self.nreads + 2) + len(self.path) + 1 \
            + (self.totlen + self.nreads + 4)
        if self.mask_anno is not None:
            # one merged track record: sizeof(HITS_TRACK)=40, name, anno, data
            nm = self.track_names[0] if len(self.track_names) == 1 else "merge"
            s += 40 + len(nm) + 1 + 8 * (self.nreads + 1) + 4 * int(self.mask_anno[self.nreads])
        return s


def _parse_name(path: str) -> Tuple[str, str, int]:
    """Split 'dir/G.3[.db]' into (dir, root, part) like Open_DB_Block (db/DB.c:473-503)."""
    d, base = os.path.split(path)
    if base.endswith(".db") or base.endswith(".dam"):
        base = base[: base.rindex(".")]
    part = 0
    m = re.match(r"^(.*)\.(\d+)$", base)
    if m and int(m.group(2)) > 0:
        base, part = m.group(1), int(m.group(2))
    return d or ".", base, part


def read_stub(dbfile: str):
    """Parse the text stub: returns (nfiles, nblocks, size, firsts[nblocks+1])."""
    with open(dbfile, "r") as f:
        toks = f.read().split("\n")
    nfiles = int(toks[0].split("=")[1])
    i = 1 + nfiles
    nblocks, size, firsts = 0, 0, []
    if i < len(toks) and toks[i].startswith("blocks"):
        nblocks = int(toks[i].split("=")[1])
        size = int(toks[i + 1].split("=")[1])
        firsts = [int(x) for x in toks[i + 2: i + 3 + nblocks] if x.strip() != ""]
    return nfiles, nblocks, size, firsts


def unpack_2bit(packed: np.ndarray, n: int) -> np.ndarray:
    """2 bit/base, 4 bases/byte, MSB first (db/DB.c:372-379)."""
    b = packed.astype(np.uint8)
    out = np.empty((b.size, 4), dtype=np.uint8)
    out[:, 0] = (b >> 6) & 3
    out[:, 1] = (b >> 4) & 3
    out[:, 2] = (b >> 2) & 3
    out[:, 3] = b & 3
    return out.reshape(-1)[:n]


def read_block(path: str, tracks: Optional[List[str]] = None) -> Block:
    d, root, part = _parse_name(path)
    stub = os.path.join(d, root + ".db")
    if not os.path.exists(stub):
        stub = os.path.join(d, root + ".dam")
    _, nblocks, _, firsts = read_stub(stub)
    hidden = os.path.join(d, "." + root)
    with open(hidden + ".idx", "rb") as f:
        hdr = f.read(HITS_DB_BYTES)
        ureads = struct.unpack_from("<i", hdr, 0)[0]
        freq = np.frombuffer(hdr, dtype="<f4", count=4, offset=4).copy()
        maxlen = struct.unpack_from("<i", hdr, 20)[0]
        totlen = struct.unpack_from("<q", hdr, 24)[0]
        if part > 0:
            if part > nblocks:
                raise ValueError(f"DB {root} has only {nblocks} blocks")
            ufirst, ulast = firsts[part - 1], firsts[part]
        else:
            ufirst, ulast = 0, ureads
        f.seek(HITS_DB_BYTES + HITS_READ_BYTES * ufirst)
        recs = np.frombuffer(f.read(HITS_READ_BYTES * (ulast - ufirst)), dtype=_READ_DT)
    nreads = ulast - ufirst
    rlen = recs["rlen"].astype(np.int32)
    if part > 0:
        maxlen = int(rlen.max()) if nreads else 0
        totlen = int(rlen.astype(np.int64).sum())
    boff = np.zeros(nreads + 1, dtype=np.int64)
    np.cumsum(rlen.astype(np.int64) + 1, out=boff[1:])
    bases = np.full(int(boff[nreads]) + 1, 4, dtype=np.uint8)
    bps = np.memmap(hidden + ".bps", dtype=np.uint8, mode="r")
    disk = recs["boff"].astype(np.int64)
    if nreads:
        # the reads of a block are contiguous in .bps, each starting on a byte: unpack the whole range in one go, then
        # copy every read's bases to its slot (unpacking read by read cost 0.25 s per 70 Mbp block)
        lo, hi = int(disk.min()), int((disk + ((rlen.astype(np.int64) + 3) >> 2)).max())
        flat = unpack_2bit(np.asarray(bps[lo:hi]), 4 * (hi - lo))
        for i in range(nreads):
            n = int(rlen[i])
            o, q = 1 + int(boff[i]), 4 * (int(disk[i]) - lo)
            bases[o: o + n] = flat[q: q + n]
    blk = Block(name=f"{root}.{part}" if part > 0 else root, root=root, path=hidden, part=part,
                ufirst=ufirst, nreads=nreads, maxlen=maxlen, totlen=totlen, ureads=ureads, freq=freq,
                rlen=rlen, boff=boff, bases=bases)
    if tracks:
        load_mask_tracks(blk, tracks)
    return blk


MARVEL_CHUNK = 8 * 1024 * 1024        # COMPRESS_MAX_CHUNK, lib/compression.c:13


def _uncompress_chunks(buf: bytes) -> bytes:
    """uncompress_chunks (lib/compression.c:79-110): a sequence of {uint64 clen, zlib stream of clen bytes}."""
    import zlib
    out, i = [], 0
    while i != len(buf):
        clen = struct.unpack_from("<Q", buf, i)[0]
        out.append(zlib.decompress(buf[i + 8: i + 8 + clen]))
        i += 8 + clen
    return b"".join(out)


def _compress_chunks(buf: bytes) -> bytes:
    """compress_chunks (lib/compression.c:15-77)."""
    import zlib
    out = []
    for i in range(0, len(buf), MARVEL_CHUNK):
        c = zlib.compress(buf[i: i + MARVEL_CHUNK])
        out.append(struct.pack("<Q", len(c)) + c)
    return b"".join(out)


def write_marvel_track(db_path: str, name: str, anno: np.ndarray, data: np.ndarray) -> None:
    """A MARVEL track of the WHOLE DB (lib/tracks.h:43-60, track_write): `<path>.<name>.a2` = 64-byte header
    {u16 version 2, u16 size 8, u32 pad, u64 len, u64 clen, u64 cdlen, 4 x u64 reserved} + the chunk-compressed byte
    offsets (uint64[len + 1]), `<path>.<name>.d2` = the chunk-compressed int32 data.  Used by the tests to make fixtures the
    reference's track_load reads."""
    anno = np.ascontiguousarray(anno, dtype="<u8")
    canno, cdata = _compress_chunks(anno.tobytes()), _compress_chunks(np.ascontiguousarray(data, dtype="<i4").tobytes())
    with open(f"{db_path}.{name}.a2", "wb") as f:
        f.write(struct.pack("<HHIQQQ4Q", 2, 8, 0, len(anno) - 1, len(canno), len(cdata), 0, 0, 0, 0))
        f.write(canno)
    with open(f"{db_path}.{name}.d2", "wb") as f:
        f.write(cdata)


def _load_marvel_track(blk: Block, name: str) -> Optional[Tuple[np.ndarray, np.ndarray]]:
    """track_load (lib/tracks.c:20-126): the `.a2/.d2` pair if it exists (else None: the caller falls back to the Dazzler
    track like :26-30), cut to the block.  As in the reference the cut is made only `if (db->ufirst > 1)` (:98) -- a block
    that starts at read 1 gets the whole-DB arrays unshifted there, and so here."""
    a2 = f"{blk.path}.{name}.a2"
    if not os.path.exists(a2):
        return None
    with open(a2, "rb") as f:
        version, size, _pad, tlen, clen, cdlen = struct.unpack("<HHIQQQ", f.read(32))
        f.read(32)
        if size <= 0:
            raise ValueError(f"invalid size field in track {name}")
        if blk.part == 0 and tlen != blk.nreads:
            raise ValueError(f"invalid track length in header of track {name}")
        canno = f.read(clen)
    if size != 8:
        raise ValueError("mask tracks must have 8-byte offsets (daligner.c:480-482)")
    anno = np.frombuffer(_uncompress_chunks(canno), dtype="<u8")[: tlen + 1].astype(np.int64)
    with open(f"{blk.path}.{name}.d2", "rb") as f:
        data = _uncompress_chunks(f.read(cdlen))[: int(anno[tlen])]
    if blk.ufirst > 1:
        anno = anno[blk.ufirst: blk.ufirst + blk.nreads + 1].copy()
        start, end = int(anno[0]), int(anno[blk.nreads])
        anno -= start
        data = data[start:end]
    else:
        anno = anno[: blk.nreads + 1].copy()
    d32 = np.frombuffer(data, dtype="<i4")[: int(anno[blk.nreads]) // 4].copy()
    return anno // 4, d32


def _load_interval_track(blk: Block, name: str) -> Tuple[np.ndarray, np.ndarray]:
    """track_load (lib/tracks.c:20): the MARVEL `.a2/.d2` track if present, else the Dazzler interval track for the block
    (db/DB.c:1113-1294); offsets converted to int units (daligner.c:480-482)."""
    mt = _load_marvel_track(blk, name)
    if mt is not None:
        return mt
    cand = []
    if blk.part:
        cand.append((f"{blk.path}.{blk.part}.{name}", True))
    cand.append((f"{blk.path}.{name}", False))
    for pre, ispart in cand:
        if os.path.exists(pre + ".anno"):
            break
    else:
        raise FileNotFoundError(f"Track '{name}' does not exist")
    with open(pre + ".anno", "rb") as f:
        tracklen, size = struct.unpack("<ii", f.read(8))
        if size != 8:
            raise ValueError("mask tracks must have 8-byte offsets (daligner.c:480-482)")
        if not ispart and blk.part > 0:
            f.seek(8 + size * blk.ufirst)
        anno = np.frombuffer(f.read(8 * (blk.nreads + 1)), dtype="<i8").copy()
    off0 = int(anno[0])
    anno -= off0
    with open(pre + ".data", "rb") as f:
        f.seek(off0)
        data = np.frombuffer(f.read(int(anno[blk.nreads])), dtype="<i4").copy()
    return anno // 4, data


def merge_interval_tracks(nreads: int, tracks: List[Tuple[np.ndarray, np.ndarray]]):
    """Union of several interval tracks per read (Merge_Size/Merge_Tracks, daligner.c:263-439):
    sweep over all endpoints, emitting [beg,end) of every maximal covered stretch; intervals that
    merely touch (end == next beg) are fused, as the reference's event sweep does."""
    anno = np.zeros(nreads + 1, dtype=np.int64)
    out: List[int] = []
    for i in range(nreads):
        iv = []
        for a, dta in tracks:
            seg = dta[int(a[i]): int(a[i + 1])]
            iv.extend((int(seg[j]), int(seg[j + 1])) for j in range(0, len(seg), 2))
        iv.sort()
        cur_b, cur_e = None, None
        for b, e in iv:
            if cur_b is None:
                cur_b, cur_e = b, e
            elif b <= cur_e:
                cur_e = max(cur_e, e)
            else:
                out.extend((cur_b, cur_e))
                cur_b, cur_e = b, e
        if cur_b is not None:
            out.extend((cur_b, cur_e))
        anno[i + 1] = len(out)
    return anno, np.asarray(out, dtype=np.int32)


def load_mask_tracks(blk: Block, names: List[str]) -> None:
    loaded = [_load_interval_track(blk, n) for n in names]
    if len(loaded) == 1:
        blk.mask_anno, blk.mask_data = loaded[0]
    else:
        blk.mask_anno, blk.mask_data = merge_interval_tracks(blk.nreads, loaded)
    blk.track_names = list(names)


# --------------------------------------------------------------------------------------------
# .las
# --------------------------------------------------------------------------------------------
LAS_REC_DT = np.dtype([("tlen", "<i4"), ("diffs", "<i4"), ("abpos", "<i4"), ("bbpos", "<i4"),
                       ("aepos", "<i4"), ("bepos", "<i4"), ("flags", "<u4"), ("aread", "<i4"),
                       ("bread", "<i4"), ("pad", "<i4")])
assert LAS_REC_DT.itemsize == 40


def parse_las(buf: bytes):
    """Return (novl, tspace, records[LAS_REC_DT], traces list-of-bytes)."""
    novl, tspace = struct.unpack_from("<qi", buf, 0)
    tbytes = 1 if tspace <= 125 else 2
    recs = np.zeros(novl, dtype=LAS_REC_DT)
    traces = []
    o = 12
    for i in range(novl):
        recs[i] = np.frombuffer(buf, dtype=LAS_REC_DT, count=1, offset=o)[0]
        o += 40
        n = int(recs[i]["tlen"]) * tbytes
        traces.append(bytes(buf[o: o + n]))
        o += n
    if o != len(buf):
        raise ValueError(f".las length mismatch: parsed {o} of {len(buf)} bytes")
    return novl, tspace, recs, traces


def las_equal(x: bytes, y: bytes) -> Tuple[bool, str]:
    """Bit-exact comparison ignoring the 4 uninitialised padding bytes 36..39 of every record
    (the reference writes struct padding, align.c:3345-3346)."""
    if len(x) != len(y):
        return False, f"length {len(x)} != {len(y)}"
    if x[:12] != y[:12]:
        return False, "header differs"
    novl, tspace = struct.unpack_from("<qi", x, 0)
    tbytes = 1 if tspace <= 125 else 2
    a = np.frombuffer(x, dtype=np.uint8).copy()
    b = np.frombuffer(y, dtype=np.uint8).copy()
    o = 12
    for i in range(novl):
        tl = struct.unpack_from("<i", x, o)[0]
        if struct.unpack_from("<i", y, o)[0] != tl:
            return False, f"record {i}: tlen differs"
        a[o + 36: o + 40] = 0
        b[o + 36: o + 40] = 0
        o += 40 + tl * tbytes
    if o != len(x):
        return False, "trailing bytes"
    if not np.array_equal(a, b):
        bad = int(np.nonzero(a != b)[0][0])
        return False, f"first differing byte at offset {bad}"
    return True, "equal"
