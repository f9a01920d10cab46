"""Python host mirror of the reference's filter interface, bound to libdamar_b200.so via ctypes.

The three operators keep the reference's names and argument meaning
(dalign/filter.h:64-70):

    set_filter_params(kmer, binshift, suppress, hitmin, nthreads)   -> Set_Filter_Params
    sort_kmers(block, comp=0)                                       -> Sort_Kmers
    match_filter(ablock, bblock, atable, btable, self_, comp)       -> Match_Filter

and ``daligner(...)`` replays daligner.c:958-1073 on top of them (index A once, compare against
every B block in both orientations, Write_Overlap_Buffer per block pair).  All arithmetic happens
in the CUDA library; this file only moves host buffers.  There is no CPU fallback: importing works
anywhere, calling any operator without the built library or without a GPU raises.
"""
from __future__ import annotations

import ctypes as C
import os
import threading
from concurrent.futures import ThreadPoolExecutor
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from .dbio import Block

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdamar_b200.so")


class BlockDesc(C.Structure):
    _fields_ = [("nreads", C.c_int32), ("maxlen", C.c_int32), ("totlen", C.c_int64), ("ufirst", C.c_int32),
                ("part", C.c_int32), ("bases", C.c_void_p), ("boff", C.c_void_p), ("rlen", C.c_void_p),
                ("mask_anno", C.c_void_p), ("mask_data", C.c_void_p), ("freq", C.c_float * 4),
                ("sizeof_db", C.c_int64), ("flags", C.c_int32), ("reserved", C.c_int32)]


class BlockView(C.Structure):
    """damar_block_view: a resident block as device pointers (device-to-device transfer between ranks)."""
    _fields_ = [("nreads", C.c_int32), ("maxlen", C.c_int32), ("ufirst", C.c_int32), ("part", C.c_int32),
                ("totlen", C.c_int64), ("nbases", C.c_int64), ("sizeof_db", C.c_int64), ("freq", C.c_float * 4),
                ("npacked_words", C.c_int64), ("packed", C.c_void_p), ("boff", C.c_void_p), ("rlen", C.c_void_p),
                ("mask_len", C.c_int64), ("mask_anno", C.c_void_p), ("mask_data", C.c_void_p),
                ("flags", C.c_int32), ("reserved", C.c_int32)]


class Options(C.Structure):
    _fields_ = [("verbose", C.c_int32), ("minover", C.c_int32), ("hgap_min", C.c_int32), ("symmetric", C.c_int32),
                ("identity", C.c_int32), ("trace_spacing", C.c_int32), ("no_trace", C.c_int32),
                ("reserved", C.c_int32), ("ave_corr", C.c_double), ("freq", C.c_float * 4),
                ("mem_limit", C.c_uint64)]


class Overlaps(C.Structure):
    _fields_ = [("novl", C.c_int64), ("ovl", C.c_void_p), ("trace_off", C.c_void_p), ("seq", C.c_void_p),
                ("trace", C.c_void_p), ("ntrace", C.c_int64), ("trace_bytes", C.c_int32), ("reserved", C.c_int32),
                ("nhits", C.c_int64), ("nfilt", C.c_int64), ("ncheck", C.c_int64),
                ("limit", C.c_int32), ("multi_la_pairs", C.c_int32)]


class Timings(C.Structure):
    _fields_ = [("upload_ms", C.c_float), ("extract_ms", C.c_float), ("ksort_ms", C.c_float),
                ("count_ms", C.c_float), ("merge_ms", C.c_float), ("hsort_ms", C.c_float),
                ("pairs_ms", C.c_float), ("align_ms", C.c_float), ("download_ms", C.c_float),
                ("kmers", C.c_int64), ("hits", C.c_int64), ("pairs", C.c_int64), ("la_calls", C.c_int64),
                ("wave_cells", C.c_int64), ("kernel_launches", C.c_int64),
                ("scatter_ms", C.c_double), ("scatter_elems", C.c_int64), ("scatter_launches", C.c_int64),
                ("align_kernel_ms", C.c_double), ("align_launches", C.c_int64), ("redo_pairs", C.c_int64),
                ("redo_kernel_ms", C.c_double), ("scatter_bytes", C.c_int64), ("wide_passes", C.c_int64),
                ("sort_ms", C.c_double), ("sort_contract_bytes", C.c_int64), ("sort_elems", C.c_int64),
                ("cells_total", C.c_int64), ("wide_total", C.c_int64), ("la_total", C.c_int64)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


KMER_DT = np.dtype([("code", "<u8"), ("rpos", "<i4"), ("read", "<i4")])
HIT_DT = np.dtype([("diag", "<i4"), ("apos", "<i4"), ("aread", "<i4"), ("bread", "<i4")])

EXPORTS = ["damar_abi_version", "damar_init", "damar_shutdown", "damar_last_error", "damar_set_filter_params",
           "damar_set_options", "damar_block_upload", "damar_block_free", "damar_sort_kmers", "damar_index_free",
           "damar_match_filter", "damar_index_download", "damar_last_hits_download", "damar_get_timings",
           "damar_sink_new", "damar_sink_free", "damar_sink_reset", "damar_sink_add", "damar_sink_count",
           "damar_sink_serialise", "damar_free", "damar_reset_kernel_timers", "damar_emission_order",
           "damar_dump_profile", "damar_worker_bind", "damar_worker_count", "damar_sink_add_at", "damar_set_biased", "damar_trace_point_diffs", "damar_las_merge",
           "damar_las_compare", "damar_sort_selftest", "damar_block_view_get", "damar_block_from_view", "damar_match_self",
           "damar_trace_point_scripts", "damar_trace_mid_scripts", "damar_host_free", "damar_dsink_new", "damar_dsink_free", "damar_dsink_reset", "damar_match_filter_sink", "damar_dsink_serialise"]

_lib = None


def load_library():
    """dlopen the in-tree CUDA library (no compute happens here)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: run `make lib` (or __graft_entry__.build()); "
                           "damar_b200 has no CPU fallback")
    L = C.CDLL(LIB_PATH)
    L.damar_abi_version.restype = C.c_int
    L.damar_init.argtypes = [C.c_int]
    L.damar_last_error.restype = C.c_char_p
    L.damar_set_filter_params.argtypes = [C.c_int] * 5
    L.damar_set_options.argtypes = [C.POINTER(Options)]
    L.damar_trace_point_diffs.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int64,
                                          C.c_int64, C.c_int, C.c_void_p, C.c_void_p]
    L.damar_trace_point_diffs.restype = C.c_int
    L.damar_trace_point_scripts.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int64,
                                            C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]
    L.damar_trace_point_scripts.restype = C.c_int
    L.damar_trace_mid_scripts.argtypes = L.damar_trace_point_scripts.argtypes
    L.damar_trace_mid_scripts.restype = C.c_int
    L.damar_host_free.argtypes = [C.c_void_p]
    L.damar_las_merge.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_int64), C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]
    L.damar_las_merge.restype = C.c_int
    L.damar_set_biased.argtypes = [C.c_int]
    L.damar_set_biased.restype = None
    L.damar_block_upload.restype = C.c_void_p
    L.damar_block_upload.argtypes = [C.POINTER(BlockDesc)]
    L.damar_block_free.argtypes = [C.c_void_p]
    L.damar_block_view_get.argtypes = [C.c_void_p, C.POINTER(BlockView)]
    L.damar_block_from_view.restype = C.c_void_p
    L.damar_block_from_view.argtypes = [C.POINTER(BlockView)]
    L.damar_sort_kmers.restype = C.c_void_p
    L.damar_sort_kmers.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int64)]
    L.damar_index_free.argtypes = [C.c_void_p]
    L.damar_match_filter.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                     C.POINTER(Overlaps)]
    L.damar_match_self.argtypes = [C.c_void_p, C.POINTER(Overlaps)]
    L.damar_dsink_new.restype = C.c_void_p
    L.damar_dsink_free.argtypes = [C.c_void_p]
    L.damar_dsink_reset.argtypes = [C.c_void_p]
    L.damar_match_filter_sink.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                          C.POINTER(Overlaps)]
    L.damar_dsink_serialise.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p),
                                        C.POINTER(C.c_int64), C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]
    L.damar_index_download.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
    L.damar_last_hits_download.restype = C.c_int64
    L.damar_last_hits_download.argtypes = [C.c_void_p, C.c_int64]
    L.damar_get_timings.argtypes = [C.POINTER(Timings)]
    L.damar_sink_new.restype = C.c_void_p
    L.damar_sink_free.argtypes = [C.c_void_p]
    L.damar_sink_reset.argtypes = [C.c_void_p]
    L.damar_sink_add.argtypes = [C.c_void_p, C.POINTER(Overlaps)]
    L.damar_sink_add_at.argtypes = [C.c_void_p, C.POINTER(Overlaps), C.c_int64]
    L.damar_worker_bind.argtypes = [C.c_int]
    L.damar_sink_count.restype = C.c_int64
    L.damar_sink_count.argtypes = [C.c_void_p]
    L.damar_sink_serialise.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p),
                                       C.POINTER(C.c_int64), C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]
    L.damar_free.argtypes = [C.c_void_p]
    L.damar_las_compare.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64]
    L.damar_las_compare.restype = C.c_int64
    L.damar_sort_selftest.argtypes = [C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_int, C.POINTER(C.c_double)]
    L.damar_sort_selftest.restype = C.c_int
    _lib = L
    return L


def init(device: int = -1) -> None:
    L = load_library()
    if L.damar_init(device) != 0:
        raise RuntimeError("damar_b200: " + L.damar_last_error().decode())


# ------------------------------------------------------------------------------------------------
@dataclass
class Settings:
    """daligner's command-line state (daligner.c:676-703) -- defaults are the reference's."""
    kmer: int = 14            # -k
    binshift: int = 6         # -w
    hitmin: int = 35          # -h
    suppress: int = 0         # -t
    ave_corr: float = 0.70    # -e
    minlen: int = 1000        # -l  (MINOVER = 2*minlen)
    trace_spacing: int = 100  # -s
    nthreads: int = 4         # -j
    mem_limit_gb: Optional[int] = None   # -M in GB; None = physical RAM like daligner.c:702-703 (the -M rule then caps k-mer
                                         # groups at MAXGRAM = 10000 mutual matches, filter.c:2634-2663); 0 = -M0: no cap
    symmetric: bool = True    # not -A
    identity: bool = False    # -I
    only_identity: bool = False  # -O
    no_trace: bool = False    # -T
    verbose: bool = False     # -v
    biased: bool = False      # -b  (variable-length k-mers by base log-frequency, filter.c:549)
    run_id: int = 1           # -r


class DeviceBlockHandle:
    """A DB block resident in HBM (2-bit packed) plus the k-mer indexes built from it.

    An index is built once per (orientation, k, -t) while the handle lives and shared by every comparison of a plan that
    needs it (daligner.c rebuilds the B index in every job, 20 sorts for the 8 distinct block orientations of a 4-block plan);
    drop_indexes() forgets them (bench.py does that at the start of every timed step, so a step always includes its sorts)."""

    def __init__(self, blk: Optional[Block], handle: Optional[int] = None):
        L = load_library()
        self.blk = blk
        self._idx: Dict[tuple, "Future"] = {}
        self._idx_lock = threading.Lock()
        if handle is not None:                               # adopted (damar_block_from_view)
            self.handle, self._keep = handle, None
            return
        self._keep = (np.ascontiguousarray(blk.bases, np.uint8), np.ascontiguousarray(blk.boff, np.int64),
                      np.ascontiguousarray(blk.rlen, np.int32),
                      None if blk.mask_anno is None else np.ascontiguousarray(blk.mask_anno, np.int64),
                      None if blk.mask_data is None else np.ascontiguousarray(
                          blk.mask_data if blk.mask_data.size else np.zeros(1, np.int32), np.int32))
        d = BlockDesc()
        d.nreads, d.maxlen, d.totlen, d.ufirst, d.part = blk.nreads, blk.maxlen, blk.totlen, blk.ufirst, blk.part
        d.bases = self._keep[0].ctypes.data + 1            # HITS_DB.bases points past the leading 4
        d.boff = self._keep[1].ctypes.data
        d.rlen = self._keep[2].ctypes.data
        d.mask_anno = None if self._keep[3] is None else self._keep[3].ctypes.data
        d.mask_data = None if self._keep[4] is None else self._keep[4].ctypes.data
        for i in range(4):
            d.freq[i] = float(blk.freq[i])
        d.sizeof_db = blk.sizeof_db()
        self.handle = L.damar_block_upload(C.byref(d))
        self._keep = None                                   # the library has its own (packed) copy now
        if not self.handle:
            raise RuntimeError("damar_b200: " + L.damar_last_error().decode())

    def view(self) -> "BlockView":
        v = BlockView()
        if load_library().damar_block_view_get(self.handle, C.byref(v)) != 0:
            raise RuntimeError("damar_block_view_get failed")
        return v

    @classmethod
    def from_view(cls, v: "BlockView", blk: Optional[Block] = None) -> "DeviceBlockHandle":
        h = load_library().damar_block_from_view(C.byref(v))
        if not h:
            raise RuntimeError("damar_b200: " + load_library().damar_last_error().decode())
        return cls(blk, handle=h)

    def index(self, comp: int, sig: tuple = ()) -> Optional["KmerIndex"]:
        """The block's sorted k-mer list for orientation `comp` under the filter settings `sig`, built by the first caller
        (on its worker context) and shared afterwards."""
        from concurrent.futures import Future
        key = (comp,) + tuple(sig)
        with self._idx_lock:
            fut = self._idx.get(key)
            mine = fut is None
            if mine:
                fut = self._idx[key] = Future()
        if mine:
            try:
                idx = sort_kmers(self, comp)
                if idx is not None:
                    idx.shared = True
                fut.set_result(idx)
            except BaseException as e:  # noqa: BLE001
                with self._idx_lock:
                    self._idx.pop(key, None)
                fut.set_exception(e)
                raise
        return fut.result()

    def drop_indexes(self):
        with self._idx_lock:
            futs, self._idx = list(self._idx.values()), {}
        for f in futs:
            idx = f.result() if f.done() and f.exception() is None else None
            if idx is not None:
                idx.shared = False
                idx.free()

    def free(self):
        self.drop_indexes()
        if self.handle:
            load_library().damar_block_free(self.handle)
            self.handle = None


class KmerIndex:
    def __init__(self, handle, length, dblock, comp):
        self.handle, self.len, self.dblock, self.comp = handle, length, dblock, comp
        self.shared = False                                 # owned by the block handle's cache: free() is a no-op

    def download(self) -> np.ndarray:
        """Sorted list in the reference's KmerPos layout (for parity tests)."""
        out = np.zeros(self.len, KMER_DT)
        if self.len and load_library().damar_index_download(self.handle, out.ctypes.data, self.len) != 0:
            raise RuntimeError("damar_index_download failed")
        return out

    def free(self):
        if self.handle and not self.shared:
            load_library().damar_index_free(self.handle)
            self.handle = None


def set_filter_params(kmer: int, binshift: int, suppress: int, hitmin: int, nthreads: int) -> int:
    """Set_Filter_Params (filter.c:171): returns 1 for an illegal combination."""
    return load_library().damar_set_filter_params(kmer, binshift, suppress, hitmin, nthreads)


def physical_memory() -> int:
    """getMemorySize (daligner.c:149-215, the sysconf branch): MEM_LIMIT's default."""
    try:
        return int(os.sysconf("SC_PHYS_PAGES")) * int(os.sysconf("SC_PAGESIZE"))
    except (ValueError, OSError):
        return 0


def set_options(s: Settings, freq: Sequence[float]) -> None:
    o = Options()
    o.verbose = int(s.verbose)
    o.minover = 2 * s.minlen
    o.hgap_min = 0
    o.symmetric = int(s.symmetric)
    o.identity = int(s.identity or s.only_identity)
    o.trace_spacing = s.trace_spacing
    o.no_trace = int(s.no_trace)
    o.ave_corr = s.ave_corr
    for i in range(4):
        o.freq[i] = float(freq[i])
    o.mem_limit = physical_memory() if s.mem_limit_gb is None else int(s.mem_limit_gb) * (1 << 30)
    load_library().damar_set_options(C.byref(o))
    load_library().damar_set_biased(int(s.biased))


def sort_kmers(dblock: DeviceBlockHandle, comp: int = 0) -> Optional[KmerIndex]:
    """Sort_Kmers (filter.c:753).  comp=1 indexes the reverse-complemented block on the device."""
    n = C.c_int64(0)
    L = load_library()
    h = L.damar_sort_kmers(dblock.handle, comp, C.byref(n))
    if not h:
        if n.value < 0:
            raise RuntimeError("damar_b200: " + L.damar_last_error().decode())
        return None
    return KmerIndex(h, n.value, dblock, comp)


def match_filter(ablock: DeviceBlockHandle, bblock: DeviceBlockHandle, atable: Optional[KmerIndex],
                 btable: Optional[KmerIndex], self_: bool, comp: int) -> Overlaps:
    """Match_Filter (filter.c:2519) for one orientation; the result lives in the library until the next call."""
    out = Overlaps()
    rc = load_library().damar_match_filter(ablock.handle, bblock.handle, atable.handle if atable else None,
                                           btable.handle if btable else None, int(self_), comp, C.byref(out))
    if rc != 0:
        raise RuntimeError("damar_b200: " + load_library().damar_last_error().decode())
    return out


def match_self(block: DeviceBlockHandle) -> Overlaps:
    """Match_Self (scrub/tandem.c:1182): every read of the block against itself; the result lives in the library until
    the next call."""
    out = Overlaps()
    if load_library().damar_match_self(block.handle, C.byref(out)) != 0:
        raise RuntimeError("damar_b200: " + load_library().damar_last_error().decode())
    return out


@dataclass
class TandemSettings:
    """datander's command-line state (scrub/datander.c:142-176) -- defaults are the reference's."""
    kmer: int = 12            # -k
    binshift: int = 4         # -w
    hitmin: int = 35          # -h
    ave_corr: float = 0.70    # -e
    minlen: int = 500         # -l  (MINOVER = 2*minlen)
    trace_spacing: int = 100  # -s
    nthreads: int = 4         # -j
    verbose: bool = False     # -v


def datander(blk: Block, s: Optional[TandemSettings] = None, dblock: Optional[DeviceBlockHandle] = None) -> Dict[str, bytes]:
    """`datander [flags] <block>` (scrub/datander.c:236-256): {"<block>.<block>.las": bytes}, the file the reference
    writes into its -o directory."""
    s = s or TandemSettings()
    init()
    if set_filter_params(s.kmer, s.binshift, 0, s.hitmin, s.nthreads):
        raise ValueError("Illegal combination of filter parameters")       # datander.c:218-222
    if blk.nreads and int(blk.rlen.min()) < s.kmer:
        raise ValueError(f"Block {blk.name} contains reads < {s.kmer}bp long !  Run DBsplit.")   # datander.c:84-88
    ds = Settings(kmer=s.kmer, binshift=s.binshift, hitmin=s.hitmin, ave_corr=s.ave_corr, minlen=s.minlen,
                  trace_spacing=s.trace_spacing, nthreads=s.nthreads, mem_limit_gb=0, verbose=s.verbose)
    set_options(ds, blk.freq)
    own = dblock is None
    db = DeviceBlockHandle(blk) if own else dblock
    try:
        ov = match_self(db)
        sink = LasSink()
        sink.add(ov, 0)
        b1, _ = sink.serialise(ds, -1)
        sink.free()
        return {f"{blk.name}.{blk.name}.las": b1.tobytes()}
    finally:
        if own:
            db.free()


def last_hits() -> np.ndarray:
    L = load_library()
    n = L.damar_last_hits_download(None, 0)
    out = np.zeros(n, HIT_DT)
    if n:
        L.damar_last_hits_download(out.ctypes.data, n)
    return out


def reset_kernel_timers() -> None:
    load_library().damar_reset_kernel_timers()


def timings() -> dict:
    t = Timings()
    load_library().damar_get_timings(C.byref(t))
    return t.as_dict()


class LasImage:
    """A serialised .las file held in library memory (no Python-side copy until asked)."""

    def __init__(self, ptr: int, nbytes: int):
        self.ptr, self.nbytes = ptr, nbytes

    @property
    def novl(self) -> int:
        return C.c_int64.from_address(self.ptr).value

    def view(self) -> np.ndarray:
        return np.ctypeslib.as_array(C.cast(self.ptr, C.POINTER(C.c_uint8)), shape=(self.nbytes,))

    def tobytes(self) -> bytes:
        return C.string_at(self.ptr, self.nbytes)

    def write(self, path: str) -> None:
        with open(path, "wb") as f:
            f.write(self.view())

    def __len__(self) -> int:
        return self.nbytes

    def __del__(self):
        if getattr(self, "ptr", None):
            load_library().damar_free(self.ptr)
            self.ptr = None


class LasSink:
    """Overlap_IO_Buffer + Write_Overlap_Buffer (align.c:5969-6380)."""

    def __init__(self):
        self.h = load_library().damar_sink_new()

    def add(self, ov: Overlaps, ordinal: Optional[int] = None):
        if ordinal is None:
            load_library().damar_sink_add(self.h, C.byref(ov))
        else:
            load_library().damar_sink_add_at(self.h, C.byref(ov), ordinal)

    def reset(self):
        load_library().damar_sink_reset(self.h)

    def count(self) -> int:
        return load_library().damar_sink_count(self.h)

    def serialise(self, s: Settings, split_read: int = -1):
        """Write_Overlap_Buffer into memory: returns (LasImage, LasImage | None)."""
        L = load_library()
        b1, b2 = C.c_void_p(), C.c_void_p()
        l1, l2 = C.c_int64(0), C.c_int64(0)
        if L.damar_sink_serialise(self.h, s.trace_spacing, int(s.no_trace), int(s.only_identity), split_read,
                                  C.byref(b1), C.byref(l1), C.byref(b2), C.byref(l2)) != 0:
            raise RuntimeError("damar_b200: " + L.damar_last_error().decode())
        o1 = LasImage(b1.value, l1.value)
        o2 = LasImage(b2.value, l2.value) if b2.value else None
        return o1, o2

    def free(self):
        if self.h:
            load_library().damar_sink_free(self.h)
            self.h = None


class DeviceSink:
    """The records of one B block's Match_Filter calls kept in HBM; serialise() sorts them (SORT_OVL) and builds the .las
    image(s) on the device and copies each finished image to the host once (align.c:6104-6367 on the device)."""

    def __init__(self):
        self.h = load_library().damar_dsink_new()

    def match_filter(self, ablock, bblock, atable, btable, self_: bool, comp: int, ordinal: int) -> Overlaps:
        out = Overlaps()
        rc = load_library().damar_match_filter_sink(ablock.handle, bblock.handle, atable.handle if atable else None,
                                                    btable.handle if btable else None, int(self_), comp, self.h, ordinal, C.byref(out))
        if rc != 0:
            raise RuntimeError("damar_b200: " + load_library().damar_last_error().decode())
        return out

    def reset(self):
        load_library().damar_dsink_reset(self.h)

    def serialise(self, s: "Settings", split_read: int, maxlen: int, maxread: int):
        L = load_library()
        b1, b2 = C.c_void_p(), C.c_void_p()
        l1, l2 = C.c_int64(0), C.c_int64(0)
        if L.damar_dsink_serialise(self.h, s.trace_spacing, int(s.no_trace), int(s.only_identity), split_read, maxlen, maxread,
                                   C.byref(b1), C.byref(l1), C.byref(b2), C.byref(l2)) != 0:
            raise RuntimeError("damar_b200: " + L.damar_last_error().decode())
        return LasImage(b1.value, l1.value), (LasImage(b2.value, l2.value) if b2.value else None)

    def free(self):
        if self.h:
            load_library().damar_dsink_free(self.h)
            self.h = None


def device_sinks() -> bool:
    """DAMAR_DEVICE_SINK=0 keeps the round-1 output path (records shipped to the host, sorted and serialised there)."""
    return os.environ.get("DAMAR_DEVICE_SINK", "1") != "0"


# ---- worker pool: independent comparisons (B block x orientation) in flight at once ----------------
# Every pool thread binds its own library context (stream, device-memory cache, result slots), so the serial
# tail of one comparison (a few long wave passes, later lasta rounds) overlaps the bulk kernels of the next and
# the host-side .las sort/serialisation overlaps GPU work.  DAMAR_WORKERS=1 runs everything on the caller's thread.
_pool = None
_pool_lock = threading.Lock()
_pool_next = [0]
_free_sinks: List["LasSink"] = []
_free_dsinks: List["DeviceSink"] = []


def _bind_worker():
    L = load_library()
    with _pool_lock:
        _pool_next[0] += 1
        idx = _pool_next[0]
    if L.damar_worker_bind(idx) != 0:
        raise RuntimeError("damar_b200: " + L.damar_last_error().decode())


def workers() -> int:
    n = int(os.environ.get("DAMAR_WORKERS", "6"))
    return max(1, min(n, load_library().damar_worker_count() - 1))


def _get_pool():
    global _pool
    if _pool is None:
        _pool = ThreadPoolExecutor(max_workers=workers(), thread_name_prefix="damar", initializer=_bind_worker)
    return _pool


class _Job:
    """One `daligner A B1 B2 ...` job between submit() and collect()."""
    __slots__ = ("ablk", "s", "symmetric", "da", "aidx", "cmps", "own", "dcache")


def _job_options(ablk: Block, bblks: List[Block], s: Settings):
    symmetric = s.symmetric and all(b.root == ablk.root for b in bblks)     # daligner.c:911-946
    return Settings(**{**s.__dict__, "symmetric": symmetric}), tuple(float(x) for x in ablk.freq)


def submit(ablk: Block, bblks: List[Block], s: Settings, dcache: Optional[Dict[str, DeviceBlockHandle]] = None) -> _Job:
    """First half of a job (daligner.c:958-1060): ships the blocks, sorts the A k-mers and queues every Match_Filter
    call (B block x orientation) on the worker pool.  collect() waits for them and writes the .las images."""
    init()
    if set_filter_params(s.kmer, s.binshift, s.suppress, s.hitmin, s.nthreads):
        raise ValueError("Illegal combination of filter parameters")       # daligner.c:862-866
    for blk in [ablk] + list(bblks):                                        # daligner.c:499-504 (read_DB)
        if blk.nreads and int(blk.rlen.min()) < s.kmer:
            raise ValueError(f"Block {blk.name} contains reads < {s.kmer}bp long !  Run DBsplit.")
    s, freq = _job_options(ablk, bblks, s)
    set_options(s, freq)
    J = _Job()
    J.ablk, J.s, J.symmetric = ablk, s, s.symmetric
    J.own = dcache is None
    J.dcache = dcache = {} if dcache is None else dcache

    def resident(b: Block) -> DeviceBlockHandle:
        if b.name not in dcache:
            dcache[b.name] = DeviceBlockHandle(b)
        return dcache[b.name]

    J.da = da = resident(ablk)
    # -b fixes its base weights at the first Sort_Kmers of a job (filter.c:774): those lists are rebuilt per job like the reference does
    sig = (s.kmer, s.suppress)
    index_of = (lambda h, comp: sort_kmers(h, comp)) if s.biased else (lambda h, comp: h.index(comp, sig))
    # The A index: with -b it belongs to the job and is built here; otherwise it is the handle's shared index, built by the
    # first comparison that asks for it on ITS worker context, so that submit() returns at once and the caller can queue
    # the next job while the pool is still busy with this one
    J.aidx = aidx = index_of(da, 0) if s.biased else None
    a_index = (lambda: aidx) if s.biased else (lambda: da.index(0, sig))
    inline = workers() == 1

    def write_buffer(bblk, selfcmp, sink):
        """Write_Overlap_Buffer for one B block (align.c:6166-6367): {file name: LasImage}"""
        an, bn = ablk.name, bblk.name
        dev = isinstance(sink, DeviceSink)
        dims = (max(ablk.maxlen, bblk.maxlen), max(ablk.ufirst + ablk.nreads, bblk.ufirst + bblk.nreads)) if dev else ()
        files = {}
        if selfcmp or not s.symmetric:                                      # align.c:6235
            b1, _ = sink.serialise(s, -1, *dims)
            files[f"{an}.{bn}.las"] = b1
        else:
            last = (bblk.ufirst + bblk.nreads - 1) if bblk.part < ablk.part else (ablk.ufirst + ablk.nreads - 1)
            b1, b2 = sink.serialise(s, last, *dims)
            n1, n2 = f"{an}.{bn}.las", f"{bn}.{an}.las"
            if bblk.part < ablk.part:                                       # align.c:6308-6313
                n1, n2 = n2, n1
            files[n1], files[n2] = b1, b2
        return files

    def compare(bblk, db, selfcmp, comp, sink, st, box):
        """one Match_Filter call (daligner.c:1041 / 1053) on the calling thread's worker context; the call that finishes a B
        block's pair of comparisons also writes the block's .las images, on its own worker context -- the sort and the image
        build are a chain of small kernels that would otherwise queue, one after the other, behind the long alignment kernels
        of the comparisons in flight while the collecting thread waits"""
        ai = a_index()
        bidx = ai if (selfcmp and comp == 0) else index_of(db, comp)
        if isinstance(sink, DeviceSink):
            ov = sink.match_filter(da, db, ai, bidx, selfcmp, comp, comp)
        else:
            ov = match_filter(da, db, ai, bidx, selfcmp, comp)
            sink.add(ov, comp)
        st[f"nhits{comp}"], st[f"nfilt{comp}"], st[f"novl{comp}"] = ov.nhits, ov.nfilt, ov.novl
        st[f"multi{comp}"] = ov.multi_la_pairs
        st[f"t{comp}"] = timings()
        if bidx is not None and bidx is not ai:
            bidx.free()
        with box["lock"]:
            box["done"] += 1
            last = box["done"] == 2
        if last:
            box["files"] = write_buffer(bblk, selfcmp, sink)

    J.cmps = []
    for bblk in bblks:
        selfcmp = bblk.name == ablk.name
        db = da if selfcmp else resident(bblk)
        if device_sinks():
            sink = _free_dsinks.pop() if _free_dsinks else DeviceSink()
        else:
            sink = _free_sinks.pop() if _free_sinks else LasSink()   # kept for the process: host buffers stay warm
        sink.reset()
        st = {}
        box = {"lock": threading.Lock(), "done": 0, "files": None}
        if inline:
            futs = None
            for comp in (0, 1):
                compare(bblk, db, selfcmp, comp, sink, st, box)
        else:
            futs = [_get_pool().submit(compare, bblk, db, selfcmp, comp, sink, st, box) for comp in (0, 1)]
        J.cmps.append((bblk, selfcmp, sink, st, futs, box))
    return J


def collect(J: _Job, stats: Optional[list] = None, as_bytes: bool = True) -> Dict[str, bytes]:
    """Second half of a job: waits for the comparisons and hands out the .las images they wrote."""
    out: Dict[str, bytes] = {}
    conv = (lambda im: im.tobytes()) if as_bytes else (lambda im: im)
    for bblk, selfcmp, sink, st, futs, box in J.cmps:
        if futs is not None:
            for f in futs:
                f.result()
        if stats is not None:
            stats.append(st)
        for name, im in box["files"].items():
            out[name] = conv(im)
        sink.reset()
        (_free_dsinks if isinstance(sink, DeviceSink) else _free_sinks).append(sink)
    if J.aidx is not None:
        J.aidx.free()
    if J.own:
        for h in J.dcache.values():
            h.free()
    return out


def daligner(ablk: Block, bblks: List[Block], s: Settings, stats: Optional[list] = None,
             dcache: Optional[Dict[str, DeviceBlockHandle]] = None, as_bytes: bool = True) -> Dict[str, bytes]:
    """`daligner [flags] A B1 B2 ...` (daligner.c:958-1073): returns {las file name: bytes}
    (or LasImage objects with as_bytes=False, which avoids one host copy per file)."""
    return collect(submit(ablk, bblks, s, dcache), stats, as_bytes)


def plan_iter(jobs, s: Settings, stats: Optional[list] = None, dcache: Optional[Dict[str, DeviceBlockHandle]] = None,
              as_bytes: bool = True):
    """Generator over the .las files of a list of jobs [(A block, [B blocks]), ...] run one after the other like the
    HPCdaligner script would (dalign/HPCdaligner.c:709-788): yields (file name, image) as jobs complete, in job order.
    Later jobs are submitted (blocks shipped, comparisons queued) BEFORE the earlier ones are collected, so the pool always
    has comparisons waiting (a plan starts with its smallest jobs, `A.1 A.1` is one block pair): up to `DAMAR_LOOKAHEAD`
    (default workers + 2) comparisons queued or running beyond the job being collected -- measured on the C2 plan: 1 -> 585-685
    ms per step, 8 -> 585-594, 24 -> 591-601 (results wait in HBM until their job is collected, so deeper is not better).  Jobs whose process-wide options
    differ (base frequencies of another DB, -b) are not overlapped."""
    dcache = {} if dcache is None else dcache
    ahead = int(os.environ.get("DAMAR_LOOKAHEAD", str(workers() + 2)))
    pend: List[_Job] = []
    pend_key = None

    def unfinished() -> int:
        return sum(1 for J in pend for c in J.cmps if c[4] is not None for f in c[4] if not f.done())

    for ablk, bblks in jobs:
        key = _job_options(ablk, bblks, s)
        if pend and (key != pend_key or s.biased):
            while pend:
                yield from collect(pend.pop(0), stats, as_bytes).items()
        while len(pend) > 1 and unfinished() >= ahead:
            yield from collect(pend.pop(0), stats, as_bytes).items()
        pend.append(submit(ablk, bblks, s, dcache))
        pend_key = key
    while pend:
        yield from collect(pend.pop(0), stats, as_bytes).items()


def daligner_plan(jobs: List[tuple], s: Settings, stats: Optional[list] = None,
                  dcache: Optional[Dict[str, DeviceBlockHandle]] = None, as_bytes: bool = True) -> Dict[str, bytes]:
    """The jobs of an HPCdaligner plan through plan_iter(): {las file name: image} of the whole plan."""
    own = dcache is None
    dcache = {} if dcache is None else dcache
    out: Dict[str, bytes] = dict(plan_iter(jobs, s, stats, dcache, as_bytes))
    if own:
        for h in dcache.values():
            h.free()
    return out


def trace_point_diffs(ablock: DeviceBlockHandle, bblock: DeviceBlockHandle, las: bytes, segments: bool = False):
    """Compute_Trace_PTS (align.c:5577-5692) in its cost form over a whole .las image whose A reads lie in `ablock` and B
    reads in `bblock`: returns the per-record sums of the exact segment edit distances (what the reference leaves in
    path->diffs; -1 for a rejected record) and, with segments=True, also one value per trace-point segment plus the
    offset of every record's first segment."""
    from . import dbio
    novl, tspace, recs, traces = dbio.parse_las(las)
    tbytes = 1 if tspace <= 125 else 2
    toff = np.zeros(novl + 1, np.uint64)
    toff[1:] = np.cumsum(recs["tlen"].astype(np.uint64))
    trace = np.frombuffer(b"".join(traces), dtype=np.uint8).copy() if novl else np.zeros(1, np.uint8)
    if trace.size == 0:
        trace = np.zeros(1, np.uint8)
    recs = np.ascontiguousarray(recs)
    rec = np.zeros(max(novl, 1), np.int32)
    ntrace = int(toff[novl])
    seg = np.zeros(max(ntrace // 2, 1), np.uint16) if segments else None
    rc = load_library().damar_trace_point_diffs(ablock.handle, bblock.handle, recs.ctypes.data, toff.ctypes.data, trace.ctypes.data,
                                                tbytes, novl, ntrace, tspace, rec.ctypes.data,
                                                seg.ctypes.data if seg is not None else None)
    if rc not in (0, 1):
        raise RuntimeError("damar_b200: " + load_library().damar_last_error().decode())
    rec = rec[:novl]
    return (rec, seg[: ntrace // 2], (toff[:novl] // 2).astype(np.int64)) if segments else rec


def trace_point_scripts(ablock: DeviceBlockHandle, bblock: DeviceBlockHandle, las: bytes, middles: bool = False):
    """Compute_Trace_PTS (align.c:5577-5692, mode GREEDIEST) over a whole .las image: per record the exact diffs AND the edit
    script the reference leaves in path->trace.  middles=True: Compute_Trace_MID (align.c:5694-5830), the flavour that aligns
    between the middles of successive trace-point segments (corrector/LAcorrect.c:545).
    Returns (rec_diffs int32[novl], script_off int64[novl + 1], script int32[])."""
    from . import dbio
    novl, tspace, recs, traces = dbio.parse_las(las)
    tbytes = 1 if tspace <= 125 else 2
    toff = np.zeros(novl + 1, np.uint64)
    toff[1:] = np.cumsum(recs["tlen"].astype(np.uint64))
    trace = np.frombuffer(b"".join(traces), dtype=np.uint8).copy() if novl else np.zeros(1, np.uint8)
    if trace.size == 0:
        trace = np.zeros(1, np.uint8)
    recs = np.ascontiguousarray(recs)
    rec = np.zeros(max(novl, 1), np.int32)
    soff = np.zeros(novl + 1, np.int64)
    sp, sl = C.c_void_p(), C.c_int64(0)
    L = load_library()
    fn = L.damar_trace_mid_scripts if middles else L.damar_trace_point_scripts
    rc = fn(ablock.handle, bblock.handle, recs.ctypes.data, toff.ctypes.data, trace.ctypes.data, tbytes, novl,
            int(toff[novl]), tspace, rec.ctypes.data, soff.ctypes.data, C.byref(sp), C.byref(sl))
    if rc not in (0, 1):
        raise RuntimeError("damar_b200: " + L.damar_last_error().decode())
    try:
        script = np.ctypeslib.as_array(C.cast(sp, C.POINTER(C.c_int32)), shape=(max(sl.value, 1),))[: sl.value].copy() if sp.value else np.zeros(0, np.int32)
    finally:
        if sp.value:
            L.damar_host_free(sp)
    return rec[:novl], soff, script


def las_merge(images: Sequence[bytes]) -> bytes:
    """LAmerge (utils/LAmerge.c): the sorted .las images of one block's pair files -> the block's single .las image."""
    L = load_library()
    n = len(images)
    bufs = (C.c_char_p * n)(*[bytes(b) for b in images])
    lens = (C.c_int64 * n)(*[len(b) for b in images])
    out, outlen = C.c_void_p(), C.c_int64()
    rc = L.damar_las_merge(n, bufs, lens, C.byref(out), C.byref(outlen))
    if rc != 0:
        raise ValueError({1: "malformed .las image", 2: "the files have different trace spacings"}.get(rc, f"damar_las_merge: {rc}"))
    try:
        return C.string_at(out.value, outlen.value)
    finally:
        L.damar_free(out)


def sort_selftest(n: int, lo_bit: int, hi_bit: int, has_vals: bool, reps: int = 1, seed: int = 1, skew: bool = False) -> float:
    """Runs the device radix sort on n pseudo-random keys and checks order, stability and permutation on the device;
    returns the mean device time of one sort in ms."""
    init()
    ms = C.c_double(0)
    if load_library().damar_sort_selftest(n, lo_bit, hi_bit, int(has_vals), reps, seed, int(skew), C.byref(ms)) != 0:
        raise RuntimeError("damar_b200: " + load_library().damar_last_error().decode())
    return ms.value


def las_compare(x, y) -> int:
    """-1 when the two .las images (bytes or LasImage) are equal up to the padding bytes 36..39 of every record, -2 when
    their lengths differ, else the offset of the first differing byte."""
    def raw(v):
        if isinstance(v, LasImage):
            return v.ptr, v.nbytes, v
        b = bytes(v)
        return C.cast(C.c_char_p(b), C.c_void_p).value, len(b), b
    px, nx, kx = raw(x)
    py, ny, ky = raw(y)
    return int(load_library().damar_las_compare(px, nx, py, ny))


def las_dir(run_id: int, part: int) -> str:
    """getDir (db/DB.c:1851): d<run:03d>_<block:05d>."""
    return f"d{run_id:03d}_{part:05d}"


def write_outputs(outdir: str, files: Dict[str, bytes], run_id: int = 1) -> List[str]:
    paths = []
    for name, data in files.items():
        part = int(name.split(".")[1])
        d = os.path.join(outdir, las_dir(run_id, part))
        os.makedirs(d, exist_ok=True)
        p = os.path.join(d, name)
        if isinstance(data, LasImage):
            data.write(p)
        else:
            with open(p, "wb") as f:
                f.write(data)
        paths.append(p)
    return paths
