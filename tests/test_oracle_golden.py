"""CPU: the oracle (oracle/daligner_oracle.c) reproduces the reference daligner's .las files
bit for bit (modulo the 4 uninitialised padding bytes per record) on every golden case.
The golden files were written by the UNMODIFIED reference binary (tests/golden/make_golden.py)."""
import pytest

import oracle_ffi as O
from cases import BIAS, BIAS_CASES, CASES, EDGE, EDGE_CASES, GRAM, GRAM_CASES, LONG, LONG_CASES, TINY, golden_files, load_block, oracle_kwargs, parse_flags
from damar_b200 import dbio


@pytest.mark.parametrize("case", sorted(CASES))
def test_oracle_matches_reference(case):
    o = parse_flags(CASES[case]["flags"])
    names = CASES[case]["blocks"]
    blks = [load_block(TINY, n, tuple(o["masks"])) for n in names]
    out = O.daligner(blks[0], blks[1:], oracle_kwargs(o), only_identity=int(o["O"]))
    gold = golden_files(case)
    assert sorted(k for k in out if k != "_stats") == sorted(gold)
    for fn, ref in gold.items():
        ok, why = dbio.las_equal(ref, out[fn])
        assert ok, f"{case}/{fn}: {why}"
    # pairs with >1 surviving alignment enter Handle_Redundancies (filter.c:2442) in the reference; the
    # golden comparison above shows that it left them unchanged on these fixtures


@pytest.mark.parametrize("case", sorted(BIAS_CASES))
def test_oracle_matches_reference_biased(case):
    """-b (biased_tuple_thread, filter.c:549-688) on an AT-rich DB (simulator -b.7); the *_h90 pair shows that the
    option changes the records on this DB (the two reference outputs differ)."""
    o = parse_flags(BIAS_CASES[case]["flags"])
    blks = [load_block(BIAS, n, tuple(o["masks"])) for n in BIAS_CASES[case]["blocks"]]
    out = O.daligner(blks[0], blks[1:], oracle_kwargs(o), only_identity=int(o["O"]))
    gold = golden_files(case, BIAS)
    assert sorted(k for k in out if k != "_stats") == sorted(gold)
    for fn, ref in gold.items():
        ok, why = dbio.las_equal(ref, out[fn])
        assert ok, f"{case}/{fn}: {why}"


@pytest.mark.parametrize("case", sorted(EDGE_CASES))
def test_oracle_matches_reference_edge_cases(case):
    """Reads of exactly k and k + 1 bases, no k-mer match at all (header-only .las), nothing passing -h, and a comparison
    across two DBs (only the A side is reported, daligner.c:911-946)."""
    o = parse_flags(EDGE_CASES[case]["flags"])
    blks = [load_block(EDGE, n, ()) for n in EDGE_CASES[case]["blocks"]]
    kw = oracle_kwargs(o)
    kw["symmetric"] = int(kw["symmetric"] and all(b.root == blks[0].root for b in blks[1:]))
    out = O.daligner(blks[0], blks[1:], kw, only_identity=int(o["O"]))
    gold = golden_files(case, EDGE)
    assert sorted(k for k in out if k != "_stats") == sorted(gold)
    for fn, ref in gold.items():
        ok, why = dbio.las_equal(ref, out[fn])
        assert ok, f"{case}/{fn}: {why}"


@pytest.mark.parametrize("case", sorted(GRAM_CASES))
def test_oracle_matches_reference_on_repeats(case):
    """K-mer groups above the MAXGRAM cap under the default -M (= physical memory; dropped) and -M0 (kept), an equal-code
    run of ~6000 entries, and one read pair with 142 local alignments (tests/golden/make_gram.py).  The oracle's unfused
    alignments go through the product's Handle_Redundancies pass compiled for the host, as in test_path_fuse.py."""
    from test_path_fuse import oracle_daligner_fused
    o = parse_flags(GRAM_CASES[case]["flags"])
    blks = [load_block(GRAM, n, ()) for n in GRAM_CASES[case]["blocks"]]
    out, _ = oracle_daligner_fused(blks[0], [blks[0] if n == GRAM_CASES[case]["blocks"][0] else b for n, b in zip(GRAM_CASES[case]["blocks"][1:], blks[1:])],
                                   oracle_kwargs(o), fuse=True)
    gold = golden_files(case, GRAM)
    assert sorted(out) == sorted(gold)
    for fn, ref in gold.items():
        ok, why = dbio.las_equal(ref, out[fn])
        assert ok, f"{case}/{fn}: {why}"


@pytest.mark.parametrize("case", sorted(LONG_CASES))
def test_oracle_matches_reference_on_long_reads(case):
    """Nanopore-like reads (34-86 kbp, 15 % error, -e.65): several A-panels per pair (filter.c:2251-2266) and wave fronts
    of more than 64 diagonals (tests/golden/make_long.py)."""
    from test_path_fuse import oracle_daligner_fused
    o = parse_flags(LONG_CASES[case]["flags"])
    blk = load_block(LONG, "L.1", ())
    out, _ = oracle_daligner_fused(blk, [blk], oracle_kwargs(o), fuse=True)
    for fn, ref in golden_files(case, LONG).items():
        ok, why = dbio.las_equal(ref, out[fn])
        assert ok, f"{case}/{fn}: {why}"


def test_default_memory_cap_changes_output():
    a, b = golden_files("gram_default", GRAM), golden_files("gram_M0", GRAM)
    assert len(b["R.1.R.1.las"]) > 100 * len(a["R.1.R.1.las"])


def test_biased_option_changes_output():
    a, b = golden_files("bias_b_h90", BIAS), golden_files("bias_plain_h90", BIAS)
    assert any(a[fn] != b[fn] for fn in a)


def test_las_roundtrip_parse():
    gold = golden_files("default")
    novl, tspace, recs, traces = dbio.parse_las(gold["T.1.T.1.las"])
    assert tspace == 100 and novl == len(recs) > 0
    # Check_Trace_Points invariants (align.c:3455-3506)
    for r, t in zip(recs[:200], traces[:200]):
        assert r["tlen"] == 2 * ((r["aepos"] - 1) // tspace - r["abpos"] // tspace + 1)
        assert sum(t[1::2]) == r["bepos"] - r["bbpos"]
        assert sum(t[0::2]) == r["diffs"]
