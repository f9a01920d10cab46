"""CPU check of damar_b200/csrc/wave_core.cuh -- the scalar wave-extension code the CUDA kernels run one
task per thread -- against the oracle's Local_Alignment (oracle/daligner_oracle.c, pinned to the reference).
The header is compiled for the host by tests/wave_core_host.cpp (test infrastructure only)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle_ffi as O

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SO = os.path.join(HERE, "_build", "libwavecore_host.so")
SRC = [os.path.join(HERE, "wave_core_host.cpp"), os.path.join(ROOT, "damar_b200", "csrc", "wave_core.cuh")]


def _lib():
    if not os.path.exists(SO) or any(os.path.getmtime(SO) < os.path.getmtime(s) for s in SRC):
        os.makedirs(os.path.dirname(SO), exist_ok=True)
        subprocess.run(["g++", "-O2", "-g", "-std=c++17", "-fPIC", "-shared", "-Wall", "-o", SO, SRC[0]], check=True)
    L = C.CDLL(SO)
    L.wc_local_alignment.restype = C.c_int
    L.wc_local_alignment.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                     C.c_int, C.c_int, C.c_int, C.c_int,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.wc_check_trim_tables.restype = C.c_int
    L.wc_check_trim_tables.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
    return L


def trim_tables(ave_corr, freq=(.25, .25, .25, .25)):
    """New_Align_Spec + set_table (align.c:226-318)."""
    bias_factor = [.690, .690, .690, .690, .780, .850, .900, .933, .966, 1.000]
    match = float(np.float32(freq[0]) + np.float32(freq[3]))
    if match > .5:
        match = 1. - match
    bias = int((match + .025) * 20. - 1.)
    if match < .2:
        bias = 3
    ave_path = int(60 * (1. - bias_factor[bias] * (1. - ave_corr)))
    mscore = int(1000 * bias_factor[bias] * (1. - ave_corr))
    dscore = 1000 - mscore
    score = np.zeros(32768, np.int16)
    table = np.zeros(32768, np.int16)
    for x in range(32768):
        s = mx = 0
        for bit in range(14, -1, -1):
            if s > mx:
                mx = s
            s += mscore if (x >> bit) & 1 else -dscore
        table[x] = s - mx
        score[x] = s
    return ave_path, score, table, dscore


def mutate(rng, s, e):
    """simulator.c-like errors: insertion .733, deletion .2, substitution .067 of the error events."""
    out = []
    for c in s:
        if rng.random() < e:
            r = rng.random()
            if r < .733:
                out.append(int(rng.integers(4))); out.append(int(c))
            elif r < .933:
                continue
            else:
                out.append(int((c + 1 + rng.integers(3)) % 4))
        else:
            out.append(int(c))
    return np.array(out, np.uint8)


def run_both(L, a, b, comp, selfie, diag, anti, ts, ave_corr, tabs, ring=32):
    ave_path, score, table, dscore = tabs
    alen, blen = len(a), len(b)
    abuf = np.concatenate([[4], a, [4]]).astype(np.int8)
    bbuf = abuf if selfie else np.concatenate([[4], b, [4]]).astype(np.int8)
    p = O.make_params(tspace=ts, ave_corr=ave_corr)
    n = 4 * (max(alen, blen) // ts + 4) + 16
    oa, ob = np.zeros(6, np.int32), np.zeros(6, np.int32)
    ota, otb = np.zeros(n, np.uint16), np.zeros(n, np.uint16)
    O.lib().ora_local_alignment(abuf.ctypes.data + 1, alen, bbuf.ctypes.data + 1, blen, comp, selfie, diag, anti, C.byref(p),
                                oa.ctypes.data, ota.ctypes.data, ob.ctypes.data, otb.ctypes.data)
    wa, wb = np.zeros(6, np.int32), np.zeros(6, np.int32)
    wta, wtb = np.zeros(n, np.uint16), np.zeros(n, np.uint16)
    stats = np.zeros(4, np.int32)
    a8, b8 = np.ascontiguousarray(a, np.uint8), np.ascontiguousarray(b, np.uint8)
    st = L.wc_local_alignment(a8.ctypes.data, alen, b8.ctypes.data, blen, comp, selfie, diag, anti, ts, ave_path,
                              dscore, ring, wa.ctypes.data, wta.ctypes.data, wb.ctypes.data,
                              wtb.ctypes.data, stats.ctypes.data)
    return st, (oa, ota[:oa[5]], ob, otb[:ob[5]]), (wa, wta[:wa[5]], wb, wtb[:wb[5]])


def same(o, w):
    return all(np.array_equal(x, y) for x, y in zip(o, w))


@pytest.fixture(scope="module")
def lib():
    return _lib()


@pytest.mark.parametrize("ave_corr", [.65, .70, .75, .85, .95])
def test_small_trim_tables_equal_the_reference_tables(lib, ave_corr):
    ave_path, score, table, dscore = trim_tables(ave_corr)
    assert lib.wc_check_trim_tables(dscore, score.ctypes.data, table.ctypes.data) == 0


@pytest.mark.parametrize("nway", [2, 3, 4])
@pytest.mark.parametrize("ts,ave_corr,err", [(100, .70, .12), (100, .70, .06), (126, .75, .10), (50, .70, .12), (100, .65, .15), (100, .95, .004)])
def test_wave_core_matches_oracle_on_overlaps(lib, ts, ave_corr, err, nway):
    lib.wc_set_nway(nway)
    rng = np.random.default_rng(1234 + ts + int(err * 1000))
    tabs = trim_tables(ave_corr)
    G = rng.integers(0, 4, 40000).astype(np.uint8)
    done = wide = 0
    for trial in range(40):
        la, lb = int(rng.integers(1500, 9000)), int(rng.integers(1500, 9000))
        a0 = int(rng.integers(0, len(G) - la))
        b0 = int(np.clip(a0 + rng.integers(-lb + 600, la - 600), 0, len(G) - lb))
        a = mutate(rng, G[a0:a0 + la], err / 2)
        b = mutate(rng, G[b0:b0 + lb], err / 2)
        comp = int(rng.integers(2))
        # seeds: shared 12-mers (plus one arbitrary point)
        idx = {}
        for i in range(0, len(b) - 12):
            idx.setdefault(b[i:i + 12].tobytes(), i)
        seeds = []
        for i in range(0, len(a) - 12, 7):
            j = idx.get(a[i:i + 12].tobytes())
            if j is not None:
                seeds.append((i + 11, j + 11))
        rng.shuffle(seeds)
        seeds = seeds[:4] + [(int(rng.integers(1, len(a))), int(rng.integers(1, len(b))))]
        for (x, y) in seeds:
            st, o, w = run_both(lib, a, b, comp, 0, x - y, x + y, ts, ave_corr, tabs)
            if st in (3, 103):          # front wider than the 32-diagonal ring: handled by the 64 ring
                wide += 1
                st, o, w = run_both(lib, a, b, comp, 0, x - y, x + y, ts, ave_corr, tabs, ring=64)
            if st in (3, 103):
                continue
            assert st == 0
            assert same(o, w), (trial, x, y, comp, [list(v) for v in o], [list(v) for v in w])
            done += 1
    assert done > 100


def test_wave_core_self_alignment(lib):
    """A read against itself off the main diagonal (align.c:1949-1968: the wave may not cross diagonal 0)."""
    lib.wc_set_nway(4)
    rng = np.random.default_rng(99)
    tabs = trim_tables(.70)
    done = 0
    for trial in range(25):
        X = rng.integers(0, 4, int(rng.integers(800, 2500))).astype(np.uint8)
        gap = rng.integers(0, 4, int(rng.integers(0, 400))).astype(np.uint8)
        pre = rng.integers(0, 4, int(rng.integers(0, 60))).astype(np.uint8)     # short flanks: the passes end at the read
        post = rng.integers(0, 4, int(rng.integers(0, 60))).astype(np.uint8)    # ends before the front outgrows the ring
        a = np.concatenate([pre, mutate(rng, X, .05), gap, mutate(rng, X, .05), post]).astype(np.uint8)
        idx = {}
        for i in range(0, len(a) - 12):
            idx.setdefault(a[i:i + 12].tobytes(), []).append(i)
        seeds = [(v[-1] + 11, v[0] + 11) for v in idx.values() if len(v) > 1][:6]
        for (x, y) in seeds + [(y, x) for (x, y) in seeds[:2]]:
            st, o, w = run_both(lib, a, a, 0, 1, x - y, x + y, 100, .70, tabs)
            if st in (3, 103):
                st, o, w = run_both(lib, a, a, 0, 1, x - y, x + y, 100, .70, tabs, ring=64)
            if st in (3, 103):
                continue
            assert st == 0
            assert same(o, w), (trial, x, y)
            done += 1
    assert done > 20


def test_wave_core_read_ends_and_short_reads(lib):
    """Seeds at the very ends, tiny reads, reads shorter than a trace spacing."""
    lib.wc_set_nway(4)
    rng = np.random.default_rng(7)
    tabs = trim_tables(.70)
    done = 0
    for trial in range(60):
        n = int(rng.integers(40, 700))
        a = rng.integers(0, 4, n).astype(np.uint8)
        b = mutate(rng, a[int(rng.integers(0, n // 3)):], .08)
        if len(b) < 20:
            continue
        for (x, y) in [(0, 0), (len(a) - 1, len(b) - 1), (len(a) // 2, min(len(b) - 1, len(a) // 2)), (len(a), len(b)), (1, len(b) - 1)]:
            for comp in (0, 1):
                st, o, w = run_both(lib, a, b, comp, 0, x - y, x + y, 100, .70, tabs)
                if st in (3, 103):
                    st, o, w = run_both(lib, a, b, comp, 0, x - y, x + y, 100, .70, tabs, ring=64)
                if st in (3, 103):
                    continue
                assert st == 0
                assert same(o, w), (trial, x, y, comp, [list(v) for v in o], [list(v) for v in w])
                done += 1
    assert done > 300
