"""GPU parity: the CUDA path (through the C-ABI of include/damar_b200.h) against
  (1) the committed reference .las files of every golden case, bit for bit (padding bytes masked),
  (2) the oracle, stage by stage: sorted k-mer list (Sort_Kmers contract) and sorted seed hits.
"""
import json
import os

import numpy as np
import pytest

import oracle_ffi as O
from cases import BIAS, BIAS_CASES, CASES, EDGE, EDGE_CASES, GOLDEN, GRAM, GRAM_CASES, LONG, LONG_CASES, TAN, TAN_CASES, TINY, golden_files, load_block, oracle_kwargs, parse_flags, settings
from damar_b200 import api, dbio

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _gpu():
    api.init()
    yield


@pytest.mark.parametrize("case", sorted(CASES))
def test_las_matches_reference(case):
    o = parse_flags(CASES[case]["flags"])
    blks = [load_block(TINY, n, tuple(o["masks"])) for n in CASES[case]["blocks"]]
    stats = []
    out = api.daligner(blks[0], blks[1:], settings(o), stats=stats)
    gold = golden_files(case)
    assert sorted(out) == sorted(gold)
    for fn, ref in gold.items():
        ok, why = dbio.las_equal(ref, out[fn])
        assert ok, f"{case}/{fn}: {why} (ref {len(ref)} bytes, got {len(out[fn])})"


REP = os.path.join(GOLDEN, "rep")
with open(os.path.join(REP, "cases.json")) as _f:
    REP_CASES = json.load(_f)


@pytest.mark.parametrize("case", sorted(REP_CASES))
def test_las_matches_reference_with_redundant_alignments(case):
    """Read pairs with several local alignments: Handle_Redundancies (filter.c:1804-2077) fuses alignments that share
    a trace point, drops contained ones and bridges narrowly parallel ones (tests/golden/make_repeats.py)."""
    o = parse_flags(REP_CASES[case]["flags"])
    blks = [load_block(REP, n, ()) for n in REP_CASES[case]["blocks"]]
    stats = []
    out = api.daligner(blks[0], blks[1:], settings(o), stats=stats)
    d = os.path.join(REP, "ref", case)
    gold = {fn: open(os.path.join(d, fn), "rb").read() for fn in sorted(os.listdir(d))}
    assert sorted(out) == sorted(gold)
    assert sum(st["multi0"] + st["multi1"] for st in stats) > 0          # the post-pass had work
    for fn, ref in gold.items():
        ok, why = dbio.las_equal(ref, out[fn])
        assert ok, f"{case}/{fn}: {why} (ref {len(ref)} bytes, got {len(out[fn])})"


@pytest.mark.parametrize("case", sorted(BIAS_CASES))
def test_las_matches_reference_biased(case):
    """-b: variable-length k-mers (biased_tuple_thread, filter.c:549-688) on the AT-rich fixture."""
    o = parse_flags(BIAS_CASES[case]["flags"])
    blks = [load_block(BIAS, n, tuple(o["masks"])) for n in BIAS_CASES[case]["blocks"]]
    out = api.daligner(blks[0], blks[1:], settings(o))
    gold = golden_files(case, BIAS)
    assert sorted(out) == sorted(gold)
    for fn, ref in gold.items():
        ok, why = dbio.las_equal(ref, out[fn])
        assert ok, f"{case}/{fn}: {why} (ref {len(ref)} bytes, got {len(out[fn])})"


@pytest.mark.parametrize("case", sorted(EDGE_CASES))
def test_las_matches_reference_edge_cases(case):
    """Corner cases against the reference binary's files: reads of exactly k bases, no k-mer match (header-only .las),
    nothing passing -h, comparison across two DBs (asymmetric)."""
    o = parse_flags(EDGE_CASES[case]["flags"])
    blks = [load_block(EDGE, n, ()) for n in EDGE_CASES[case]["blocks"]]
    out = api.daligner(blks[0], blks[1:], settings(o))
    gold = golden_files(case, EDGE)
    assert sorted(out) == sorted(gold)
    for fn, ref in gold.items():
        ok, why = dbio.las_equal(ref, out[fn])
        assert ok, f"{case}/{fn}: {why} (ref {len(ref)} bytes, got {len(out[fn])})"


@pytest.mark.parametrize("case", sorted(GRAM_CASES))
def test_las_matches_reference_on_repeats(case):
    """K-mer groups of >= 10000 mutual matches (dropped under the default -M = physical memory, kept with -M0,
    filter.c:2634-2663), an equal-code run of ~6000 entries (the merge kernel's global-memory path) and ONE read pair
    with 142 local alignments (the growing per-pair interval list + Handle_Redundancies over 142 paths)."""
    o = parse_flags(GRAM_CASES[case]["flags"])
    blks = [load_block(GRAM, n, ()) for n in GRAM_CASES[case]["blocks"]]
    stats = []
    out = api.daligner(blks[0], [blks[0] if n == GRAM_CASES[case]["blocks"][0] else b for n, b in zip(GRAM_CASES[case]["blocks"][1:], blks[1:])],
                       settings(o), stats=stats)
    gold = golden_files(case, GRAM)
    assert sorted(out) == sorted(gold)
    for fn, ref in gold.items():
        ok, why = dbio.las_equal(ref, out[fn])
        assert ok, f"{case}/{fn}: {why} (ref {len(ref)} bytes, got {len(out[fn])})"
    if case == "gram_default":
        assert stats[0]["t0"]["hits"] < 100000          # the shared element's k-mers were capped ...
    if case == "gram_M0":
        assert stats[0]["t0"]["hits"] > 3000000         # ... and are all there without the cap


@pytest.mark.parametrize("case", sorted(LONG_CASES))
def test_las_matches_reference_on_long_reads(case):
    """Reads of 34-86 kbp at 15 % error with -e.65 (configs[3] in miniature, reference binary's files): every pair is cut
    into several A-panels (filter.c:2251-2266), passes outgrow the 32-diagonal ring (parked and resumed by a wide lane) and
    some the 64-diagonal one (the pair is redone by the strided warp kernel)."""
    o = parse_flags(LONG_CASES[case]["flags"])
    blk = load_block(LONG, "L.1", ())
    assert int(blk.rlen.max()) > 50000 + 10000                  # more than one panel
    stats = []
    out = api.daligner(blk, [blk], settings(o), stats=stats)
    gold = golden_files(case, LONG)
    assert sorted(out) == sorted(gold)
    for fn, ref in gold.items():
        ok, why = dbio.las_equal(ref, out[fn])
        assert ok, f"{case}/{fn}: {why} (ref {len(ref)} bytes, got {len(out[fn])})"
    wide = sum(st[f"t{c}"]["wide_passes"] for st in stats for c in (0, 1))
    redo = sum(st[f"t{c}"]["redo_pairs"] for st in stats for c in (0, 1))
    print(f"{case}: wide passes {wide}, pairs redone by the strided kernel {redo}")
    assert wide > 0


def test_midsize_block_pair_matches_reference():
    """One non-self block pair of the bench workload (BASELINE configs[1]: two 70 Mbp blocks, ~276 k overlaps in four .las
    files) against the reference daligner run on this host, byte for byte."""
    import shutil
    import subprocess
    import sys
    from cases import ROOT
    sys.path.insert(0, ROOT)
    import bench
    exe = os.path.join(ROOT, "oracle", "_ref", "daligner")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/daligner was not built (needs /root/reference at build time)")
    dbdir, root, nblocks = bench.make_db("c2")
    blks = [dbio.read_block(os.path.join(dbdir, f"{root}.{i}")) for i in (1, 2)]
    s = api.Settings(mem_limit_gb=32)
    out = api.daligner(blks[1], [blks[0]], s, as_bytes=False)
    for d in os.listdir(dbdir):
        if d.startswith("d001_"):
            shutil.rmtree(os.path.join(dbdir, d))
    subprocess.run([exe, f"-j{bench.pow2_floor(os.cpu_count() or 1)}", "-M32", f"{root}.2", f"{root}.1"], cwd=dbdir, check=True,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    n = 0
    for fn, im in out.items():
        part = int(fn.split(".")[1])
        ref = open(os.path.join(dbdir, f"d001_{part:05d}", fn), "rb").read()
        assert api.las_compare(ref, im) == -1, fn
        n += im.novl
    assert n > 200000
    for d in os.listdir(dbdir):
        if d.startswith("d001_"):
            shutil.rmtree(os.path.join(dbdir, d))


@pytest.mark.parametrize("case", sorted(TAN_CASES))
def test_datander_matches_reference(case):
    """datander (scrub/datander.c, scrub/tandem.c): every read against itself for tandem repeats -- reads with tandem arrays,
    one longer than a 50 kbp panel, against the reference datander's files (tests/golden/make_tandem.py)."""
    o = O.parse_tandem_flags(TAN_CASES[case]["flags"])
    blk = load_block(TAN, TAN_CASES[case]["blocks"][0], ())
    s = api.TandemSettings(kmer=o["k"], binshift=o["w"], hitmin=o["h"], ave_corr=o["e"], minlen=o["l"], trace_spacing=o["s"], nthreads=o["j"])
    out = api.datander(blk, s)
    d = os.path.join(TAN, "ref", case)
    gold = {fn: open(os.path.join(d, fn), "rb").read() for fn in sorted(os.listdir(d))}
    assert sorted(out) == sorted(gold)
    for fn, ref in gold.items():
        ok, why = dbio.las_equal(ref, out[fn])
        assert ok, f"{case}/{fn}: {why} (ref {len(ref)} bytes, got {len(out[fn])})"


@pytest.mark.parametrize("device_sink", ["1", "0"])
def test_only_identity_on_tandem_reads(device_sink):
    """daligner -O (ONLY_IDENTITY, daligner.c:756, align.c:6185-6194): only alignments of a read with itself are written --
    non-trivial on the reads with tandem arrays (reference daligner's file), through both output paths."""
    import subprocess
    import sys
    from cases import ROOT
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
            "from cases import *\nfrom damar_b200 import api, dbio\n"
            "blk = load_block(TAN, 'N.1', ())\n"
            "out = api.daligner(blk, [blk], api.Settings(only_identity=True, minlen=500, mem_limit_gb=0))\n"
            "ref = open(os.path.join(TAN, 'ref_dal', 'tan_dal_O', 'N.1.N.1.las'), 'rb').read()\n"
            "ok, why = dbio.las_equal(ref, out['N.1.N.1.las'])\n"
            "assert ok, why\nassert len(ref) > 10000\nprint('only identity ok')\n") % (ROOT, os.path.join(ROOT, "tests"))
    r = subprocess.run([sys.executable, "-c", code], env={**os.environ, "DAMAR_DEVICE_SINK": device_sink}, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "only identity ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize("case", ["tan_default", "tan_k14"])
def test_dropin_reference_datander_host(case, tmp_path):
    """The UNMODIFIED scrub/datander.c linked against libdamar_b200.so through the tandem.h shim (oracle/_ref/datander_b200)."""
    import shutil
    import subprocess
    from cases import ROOT
    exe = os.path.join(ROOT, "oracle", "_ref", "datander_b200")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/datander_b200 not built (needs /root/reference at build time)")
    for fn in os.listdir(TAN):
        p = os.path.join(TAN, fn)
        if os.path.isfile(p):
            shutil.copy(p, tmp_path / fn)
    r = subprocess.run([exe] + TAN_CASES[case]["flags"] + ["-o", "tan"] + TAN_CASES[case]["blocks"], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    d = os.path.join(TAN, "ref", case)
    for fn in sorted(os.listdir(d)):
        ok, why = dbio.las_equal(open(os.path.join(d, fn), "rb").read(), open(tmp_path / "tan" / fn, "rb").read())
        assert ok, f"{case}/{fn}: {why}"


@pytest.mark.parametrize("env", [{"DAMAR_PACK": "0"}, {"DAMAR_DEVICE_SINK": "0"}, {"DAMAR_MAX_HITS": "30000", "DAMAR_PROFILE": "1"},
                                 {"DAMAR_MAX_HITS": "30000", "DAMAR_DEVICE_SINK": "0", "DAMAR_PROFILE": "1"},
                                 {"DAMAR_MAX_HITS": "30000", "DAMAR_PACK": "0", "DAMAR_PROFILE": "1"}],
                         ids=["unpacked", "host_sink", "chunked", "chunked_host_sink", "chunked_unpacked"])
def test_alternative_paths_give_the_same_files(env):
    """DAMAR_PACK=0 keeps (code, gpos) and (key, bpos) as pairs instead of one 64-bit word per k-mer / seed hit -- the
    layout blocks too large for the packed words fall back to.  DAMAR_DEVICE_SINK=0 ships the records to the host and sorts
    and serialises them there (the path the drop-in shim uses) instead of building the .las image on the device.
    DAMAR_MAX_HITS lowers the number of seed hits one sort takes (2^30 by default), so that these small comparisons are cut
    into chunks of B reads the way a comparison with billions of hits is (the reference's hit array is int64-indexed,
    filter.c:2529): slice rule, emission ranks and the -M limit must come out as for the whole.  Run in a fresh process
    (the switches are read once)."""
    import subprocess
    import sys
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
            "from cases import *\nfrom damar_b200 import api, dbio\n"
            "import json, os\n"
            "REP = os.path.join(GOLDEN, 'rep'); REP_CASES = json.load(open(os.path.join(REP, 'cases.json')))\n"
            "chunked = 'DAMAR_MAX_HITS' in os.environ\n"   # one read of gram_M0 alone has 8.9 M hits (a poly-A run): no chunk of reads is that small
            "for table, dbdir, names in ((CASES, TINY, ['default', 'identity', 'rndmask', 'k16s200', 'asym', 'notrace', 'b2']), (GRAM_CASES, GRAM, ['many_la'] if chunked else ['gram_M0', 'many_la']), (BIAS_CASES, BIAS, ['bias_b']), (REP_CASES, REP, sorted(REP_CASES) if chunked else []), (LONG_CASES, LONG, sorted(LONG_CASES) if chunked else [])):\n"
            "    for case in names:\n"
            "        o = parse_flags(table[case]['flags'])\n"
            "        blks = {n: load_block(dbdir, n, tuple(o['masks'])) for n in set(table[case]['blocks'])}\n"
            "        nm = table[case]['blocks']\n"
            "        out = api.daligner(blks[nm[0]], [blks[n] for n in nm[1:]], settings(o))\n"
            "        for fn, ref in golden_files(case, dbdir).items():\n"
            "            ok, why = dbio.las_equal(ref, out[fn])\n"
            "            assert ok, (case, fn, why)\n"
            "print('unpacked ok')\n") % (os.path.dirname(GOLDEN).rsplit('/tests', 1)[0], os.path.dirname(GOLDEN))
    r = subprocess.run([sys.executable, "-c", code], env={**os.environ, **env}, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "unpacked ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
    if "DAMAR_MAX_HITS" in env:
        assert "mf.chunked" in r.stderr, "no comparison was cut into chunks"


def test_reads_shorter_than_k_are_refused():
    """daligner.c:499-504: a block with a read shorter than k stops the run before any work is done."""
    blk = load_block(EDGE, "E.1", ())
    with pytest.raises(ValueError, match="contains reads < 16bp long"):
        api.daligner(blk, [blk], api.Settings(kmer=16))


KMER_CASES = [(TINY, CASES, c, comp) for c, comp in [("default", 0), ("default", 1), ("rndmask", 0), ("rndmask", 1),
                                                     ("k16s200", 0), ("twomasks", 1)]] + \
             [(BIAS, BIAS_CASES, c, comp) for c, comp in [("bias_b", 0), ("bias_b", 1), ("bias_b_dust", 0),
                                                          ("bias_b_dust", 1), ("bias_b_k16", 1)]]


@pytest.mark.parametrize("dbdir,table,case,comp", KMER_CASES, ids=[f"{c}-{comp}" for _, _, c, comp in KMER_CASES])
def test_sorted_kmers_match_oracle(dbdir, table, case, comp):
    o = parse_flags(table[case]["flags"])
    blk = load_block(dbdir, table[case]["blocks"][0], tuple(o["masks"]))
    p = O.make_params(**{**oracle_kwargs(o), "freq": tuple(float(x) for x in blk.freq)})
    hb = O.HostBlock(blk)
    L = O.lib()
    if comp:
        import ctypes as C
        cb = L.ora_complement_block(C.byref(hb.c))
        n = C.c_int64(0)
        ptr = L.ora_sort_kmers(cb, C.byref(p), C.byref(n))
        ref = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(n.value * 16,)).view(O.KMER_DT).copy()
        L.ora_free(ptr); L.ora_free_block(cb)
    else:
        ptr, n, arr = O.sort_kmers(hb, p)
        ref = arr[:n].copy()
        L.ora_free(ptr)
    assert api.set_filter_params(o["k"], o["w"], o["t"], o["h"], o["j"]) == 0
    api.set_options(settings(o), blk.freq)
    d = api.DeviceBlockHandle(blk)
    if comp and o["b"]:
        api.sort_kmers(d, 0).free()        # -b weights are fixed by the first (forward) Sort_Kmers call of a job
    idx = api.sort_kmers(d, comp)
    got = idx.download()
    idx.free(); d.free()
    assert len(got) == len(ref)
    assert np.array_equal(got["code"], ref["code"])
    assert np.array_equal(got["read"], ref["read"])
    assert np.array_equal(got["rpos"], ref["rpos"])


@pytest.mark.parametrize("case,comp", [("default", 0), ("default", 1), ("identity", 0), ("identity", 1), ("memlimit", 0)])
def test_sorted_hits_match_oracle(case, comp):
    import ctypes as C
    o = parse_flags(CASES[case]["flags"])
    names = CASES[case]["blocks"]
    ablk = load_block(TINY, names[0], tuple(o["masks"]))
    bblk = load_block(TINY, names[1], tuple(o["masks"]))
    selfcmp = names[0] == names[1]
    kw = oracle_kwargs(o)
    kw["freq"] = tuple(float(x) for x in ablk.freq)
    p = O.make_params(**kw)
    p.sizeof_a, p.sizeof_b = ablk.sizeof_db(), bblk.sizeof_db()
    L = O.lib()
    ha = O.HostBlock(ablk)
    hb = ha if selfcmp else O.HostBlock(bblk)
    aptr, alen, _ = O.sort_kmers(ha, p)
    if comp:
        cb = L.ora_complement_block(C.byref(hb.c))
        n = C.c_int64(0)
        bptr = L.ora_sort_kmers(cb, C.byref(p), C.byref(n)); blen = n.value
        nh = C.c_int64(0); lim = C.c_int32(0)
        hp = L.ora_seed_hits(C.byref(ha.c), cb, aptr, alen, bptr, blen, int(selfcmp), 1, C.byref(p), C.byref(nh), C.byref(lim))
    else:
        if selfcmp:
            bptr, blen = aptr, alen
        else:
            bptr, blen, _ = O.sort_kmers(hb, p)
        nh = C.c_int64(0); lim = C.c_int32(0)
        hp = L.ora_seed_hits(C.byref(ha.c), C.byref(hb.c), aptr, alen, bptr, blen, int(selfcmp), 0, C.byref(p), C.byref(nh), C.byref(lim))
    ref = np.ctypeslib.as_array(C.cast(hp, C.POINTER(C.c_uint8)), shape=(nh.value * 16,)).view(O.HIT_DT).copy()

    assert api.set_filter_params(o["k"], o["w"], o["t"], o["h"], o["j"]) == 0
    api.set_options(settings(o), ablk.freq)
    da = api.DeviceBlockHandle(ablk)
    db = da if selfcmp else api.DeviceBlockHandle(bblk)
    ai = api.sort_kmers(da, 0)
    bi = ai if (selfcmp and not comp) else api.sort_kmers(db, comp)
    ov = api.match_filter(da, db, ai, bi, selfcmp, comp)
    got = api.last_hits()
    assert ov.nhits == nh.value and ov.limit == lim.value
    for f in ("bread", "aread", "apos", "diag"):
        assert np.array_equal(got[f], ref[f]), f


# ------------------------------------------------------------------------------------------------
# drop-in: the UNMODIFIED reference host (daligner.c + align.c + DB.c) linked against our library
# through the filter.h shim (oracle/_ref/daligner_b200, built by oracle/Makefile)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", ["default", "b2", "dust", "twomasks", "identity", "asym", "k16s200", "memlimit", "notrace",
                                  "bias_b", "bias_b_dust"])
def test_dropin_reference_host(case, tmp_path):
    import os, shutil, subprocess
    from cases import ROOT
    exe = os.path.join(ROOT, "oracle", "_ref", "daligner_b200")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/daligner_b200 not built (needs /root/reference at build time)")
    dbdir, table = (BIAS, BIAS_CASES) if case.startswith("bias") else (TINY, CASES)
    for fn in os.listdir(dbdir):
        p = os.path.join(dbdir, fn)
        if os.path.isfile(p):
            shutil.copy(p, tmp_path / fn)
    r = subprocess.run([exe] + table[case]["flags"] + table[case]["blocks"], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    gold = golden_files(case, dbdir)
    got = {}
    for d in os.listdir(tmp_path):
        if d.startswith("d001_"):
            for fn in os.listdir(tmp_path / d):
                got[fn] = open(tmp_path / d / fn, "rb").read()
    assert sorted(got) == sorted(gold)
    for fn, ref in gold.items():
        ok, why = dbio.las_equal(ref, got[fn])
        assert ok, f"{case}/{fn}: {why}"


def test_sharded_plan_world1(tmp_path):
    """multi.run_sharded at world size 1 = the whole HPCdaligner plan; files land in d<run>_<blk>/."""
    import os
    from damar_b200 import multi
    o = parse_flags(CASES["default"]["flags"])
    novl, files = multi.run_sharded(TINY, "T", 2, settings(o), 0, 1, outdir=str(tmp_path))
    gold = dict(golden_files("b2"))                       # job `daligner T.2 T.1 T.2`
    gold["T.1.T.1.las"] = golden_files("default")["T.1.T.1.las"]   # job `daligner T.1 T.1`
    assert sorted(files) == sorted(gold)
    for fn, ref in gold.items():
        part = int(fn.split(".")[1])
        got = open(os.path.join(tmp_path, f"d001_{part:05d}", fn), "rb").read()
        ok, why = dbio.las_equal(ref, got)
        assert ok, f"{fn}: {why}"
    assert novl == sum(int.from_bytes(v[:8], "little") for v in gold.values())


def test_pipelined_plan_matches_reference():
    """api.daligner_plan overlaps job i+1 (upload, A index, queued comparisons) with the collection of job i: same files."""
    o = parse_flags(CASES["default"]["flags"])
    b1, b2 = load_block(TINY, "T.1", ()), load_block(TINY, "T.2", ())
    out = api.daligner_plan([(b1, [b1]), (b2, [b1, b2])], settings(o))
    gold = dict(golden_files("b2"))                       # job `daligner T.2 T.1 T.2`
    gold["T.1.T.1.las"] = golden_files("default")["T.1.T.1.las"]   # job `daligner T.1 T.1`
    assert sorted(out) == sorted(gold)
    for fn, ref in gold.items():
        ok, why = dbio.las_equal(ref, out[fn])
        assert ok, f"{fn}: {why}"


def test_properties_at_scale():
    """20 Mbp / 2 blocks: size-independent invariants of the emitted .las (what LAcheck -p -s checks,
    utils/LAcheck.c:116-221, align.c:3455-3506) plus A/B symmetry, and equality with the reference
    binary when it is available on the box."""
    import os, shutil, subprocess, sys
    from cases import ROOT
    sys.path.insert(0, ROOT)
    import bench
    dbdir, root, nblocks = bench.make_db("small")
    blks = [dbio.read_block(os.path.join(dbdir, f"{root}.{i}")) for i in (1, 2)]
    s = api.Settings(mem_limit_gb=32)
    out = api.daligner(blks[1], [blks[0], blks[1]], s)
    n12 = int.from_bytes(out[f"{root}.1.{root}.2.las"][:8], "little")
    n21 = int.from_bytes(out[f"{root}.2.{root}.1.las"][:8], "little")
    assert n12 == n21 > 1000
    for fn, buf in out.items():
        novl, tspace, recs, traces = dbio.parse_las(buf)
        keys = [(int(r["aread"]), int(r["bread"]), int(r["flags"]) & 1, int(r["abpos"])) for r in recs]
        assert keys == sorted(keys), fn
        for r, t in list(zip(recs, traces))[::7]:
            assert r["tlen"] == 2 * ((r["aepos"] - 1) // tspace - r["abpos"] // tspace + 1)
            assert sum(t[1::2]) == r["bepos"] - r["bbpos"] and sum(t[0::2]) == r["diffs"]
            assert (r["aepos"] - r["abpos"]) + (r["bepos"] - r["bbpos"]) >= 2000
    # Compute_Trace_PTS on the device (trace_pts.cu): every trace-point segment the wave kernel emitted is realisable with at
    # most its recorded differences, and the per-record sums equal the records' own diffs except after bridging
    import time
    dh = {n: api.DeviceBlockHandle(blks[i]) for i, n in enumerate((f"{root}.1", f"{root}.2"))}
    for fn, buf in out.items():
        an, bn = ".".join(fn.split(".")[0:2]), ".".join(fn.split(".")[2:4])
        novl, tspace, recs, traces = dbio.parse_las(buf)
        t0 = time.perf_counter()
        rec, seg, soff = api.trace_point_diffs(dh[an], dh[bn], buf, segments=True)
        dt = time.perf_counter() - t0
        d_rec = np.concatenate([np.frombuffer(t, dtype=np.uint8)[0::2].astype(np.int64) for t in traces])
        assert np.all(rec >= 0) and np.all(seg <= d_rec), fn
        assert np.all(rec <= recs["diffs"]) and (rec == recs["diffs"]).mean() > 0.99, fn
        print(f"trace_point_diffs {fn}: {novl} records, {len(seg)} segments in {dt * 1e3:.1f} ms (incl. parse + copies)")
    for h in dh.values():
        h.free()
    exe = os.path.join(ROOT, "oracle", "_ref", "daligner")
    if os.path.exists(exe):
        for d in os.listdir(dbdir):
            if d.startswith("d001_"):
                shutil.rmtree(os.path.join(dbdir, d))
        subprocess.run([exe, "-j4", "-M32", f"{root}.2", f"{root}.1", f"{root}.2"], cwd=dbdir, check=True,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        for fn, buf in out.items():
            part = int(fn.split(".")[1])
            ref = open(os.path.join(dbdir, f"d001_{part:05d}", fn), "rb").read()
            ok, why = dbio.las_equal(ref, buf)
            assert ok, f"{fn}: {why}"


SORT_CASES = [
    # n, lo_bit, hi_bit, payload, skew
    (1, 0, 28, False, False), (2, 0, 28, True, False), (777, 0, 28, True, False),            # one partial tile
    (8192, 0, 28, False, False), (8193, 30, 58, False, False),                              # exactly one tile / one key more
    (6144 * 3 + 5, 0, 44, True, False),                                                     # pairs: 5 passes of 9 bits
    (1 << 20, 28, 56, False, False), (1 << 20, 28, 56, False, True),                        # packed k-mer records (k = 14)
    (3_000_001, 0, 43, True, True), (5_000_000, 16, 60, False, False),                      # seed-hit keys
    (1_000_000, 0, 64, True, False), (1_000_000, 3, 11, False, True),                       # all 64 bits / one narrow digit
]


@pytest.mark.parametrize("n,lo,hi,vals,skew", SORT_CASES)
def test_radix_sort_is_stable_and_sorted(n, lo, hi, vals, skew):
    """The one-sweep LSD radix sort (lex_sort's contract, filter.c:230-435: ascending on the selected bits, stable),
    checked on the device against order, stability (payload = input index / low key bits = input index) and a
    permutation checksum."""
    api.sort_selftest(n, lo, hi, vals, reps=2, seed=n, skew=skew)


@pytest.mark.parametrize("nth", [2, 15, 40, 80, 110])
def test_out_of_device_memory_is_retried_alone(nth):
    """DAMAR_FAKE_OOM=n makes the n-th device-memory request that reaches the driver fail once.  Whatever call it hits (block
    upload, Sort_Kmers, the seed stage, the alignment stage, the .las writer) is unwound, run again with the device to itself
    (`memory.solo_rerun` in the host profile) and the job's files are those of an undisturbed run -- the library neither exits
    nor returns a partial result."""
    import subprocess
    import sys
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
            "from cases import *\nfrom damar_b200 import api, dbio\n"
            "o = parse_flags(CASES['default']['flags'])\n"
            "blks = {n: load_block(TINY, n, tuple(o['masks'])) for n in set(CASES['default']['blocks'])}\n"
            "nm = CASES['default']['blocks']\n"
            "out = api.daligner(blks[nm[0]], [blks[n] for n in nm[1:]], settings(o))\n"
            "for fn, ref in golden_files('default', TINY).items():\n"
            "    ok, why = dbio.las_equal(ref, out[fn])\n"
            "    assert ok, (fn, why)\n"
            "print('files ok')\n") % (os.path.dirname(GOLDEN).rsplit('/tests', 1)[0], os.path.dirname(GOLDEN))
    r = subprocess.run([sys.executable, "-c", code], env={**os.environ, "DAMAR_FAKE_OOM": str(nth), "DAMAR_PROFILE": "1"},
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "files ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
    assert "memory.solo_rerun" in r.stderr, "the injected failure did not happen: fewer requests than expected"
