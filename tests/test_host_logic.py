"""CPU tests of the host side: DB/track/.las I/O against the reference tools' own files, the
synthetic DB generator against simulator|FA2db|DBsplit, the plan/sharding logic, the C-ABI surface,
and the multi-rank block exchange (gloo, world_size 2)."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from cases import CASES, GOLDEN, TINY, ROOT, golden_files, load_block
from damar_b200 import api, dbio, plan


def test_library_exports_every_declared_symbol():
    """Every function declared in include/*.h is exported by the built library (no compute)."""
    L = api.load_library()
    declared = set()
    for hdr in ("damar_b200.h",):
        txt = open(os.path.join(ROOT, "include", hdr)).read()
        txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
        declared |= set(re.findall(r"\b(damar_[a-z0-9_]+)\s*\(", txt))
    assert len(declared) >= 20
    missing = [s for s in sorted(declared) if not hasattr(L, s)]
    assert not missing, missing
    assert L.damar_abi_version() == 1


def test_no_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    L = api.load_library()
    assert L.damar_init(-1) != 0
    assert b"no CPU fallback" in L.damar_last_error()
    with pytest.raises(RuntimeError):
        api.init()


def test_set_filter_params_rejects_illegal():
    assert api.set_filter_params(1, 6, 0, 35, 4) == 1          # filter.c:173-174
    assert api.set_filter_params(14, 6, 0, 35, 4) == 0


def test_db_block_reader():
    b1, b2 = load_block(TINY, "T.1", ()), load_block(TINY, "T.2", ())
    assert b1.ufirst == 0 and b2.ufirst == b1.nreads
    assert b1.bases[0] == 4 and b1.bases[-1] == 4
    assert int(b1.boff[-1]) == b1.totlen + b1.nreads
    assert b1.bases.max() == 4 and set(np.unique(b1.bases[1:1 + int(b1.rlen[0])])) <= {0, 1, 2, 3}
    assert abs(float(b1.freq.sum()) - 1.0) < 1e-3
    whole = load_block(TINY, "T", ())
    assert whole.nreads == b1.nreads + b2.nreads and whole.totlen == b1.totlen + b2.totlen
    n = int(b2.rlen[0])
    o2 = 1 + int(whole.boff[b1.nreads])
    assert np.array_equal(whole.bases[o2:o2 + n], b2.bases[1:1 + n])


def test_mask_tracks_and_merge():
    d = load_block(TINY, "T.2", ("dust",))
    r = load_block(TINY, "T.2", ("rnd",))
    both = load_block(TINY, "T.2", ("rnd", "dust"))
    assert r.mask_anno[-1] == r.mask_data.size and r.mask_data.size % 2 == 0
    for blk in (d, r, both):
        for i in range(0, blk.nreads, 37):
            iv = blk.mask_data[int(blk.mask_anno[i]):int(blk.mask_anno[i + 1])]
            assert np.all(np.diff(iv) >= 0) and (iv.size == 0 or iv[-1] <= blk.rlen[i])
    # the union covers exactly the bases either track covers
    i = int(np.argmax(np.diff(r.mask_anno)))
    cover = np.zeros(int(r.rlen[i]) + 1, bool)
    for blk in (d, r):
        iv = blk.mask_data[int(blk.mask_anno[i]):int(blk.mask_anno[i + 1])]
        for j in range(0, iv.size, 2):
            cover[iv[j]:iv[j + 1]] = True
    cu = np.zeros_like(cover)
    iv = both.mask_data[int(both.mask_anno[i]):int(both.mask_anno[i + 1])]
    for j in range(0, iv.size, 2):
        cu[iv[j]:iv[j + 1]] = True
    assert np.array_equal(cover, cu)


def test_las_equal_masks_padding_only():
    g = golden_files("default")["T.1.T.1.las"]
    b = bytearray(g)
    b[12 + 36] ^= 0xFF                      # padding byte of record 0
    assert dbio.las_equal(g, bytes(b))[0]
    b[12 + 8] ^= 0x01                       # abpos of record 0
    assert not dbio.las_equal(g, bytes(b))[0]


def test_simdb_matches_reference_pipeline(tmp_path):
    """damar_b200/host/simdb == simulator | FA2db -x | DBsplit -s of the reference (bit for bit)."""
    ref = os.path.join(ROOT, "oracle", "_ref")
    if not os.path.exists(os.path.join(ref, "simulator")):
        pytest.skip("oracle/_ref not built")
    subprocess.run(["make", "-s", "-C", ROOT, "host"], check=True)
    a, b = tmp_path / "a", tmp_path / "b"
    a.mkdir(); b.mkdir()
    args = ["0.2", "-c8", "-r5", "-e.13", "-m2500", "-s700", "-x600"]
    with open(a / "sim.fasta", "w") as f:
        subprocess.run([os.path.join(ref, "simulator")] + args, stdout=f, check=True)
    subprocess.run([os.path.join(ref, "FA2db"), "-x", "700", "Q", "sim.fasta"], cwd=a, check=True)
    subprocess.run([os.path.join(ref, "DBsplit"), "-s1", "Q"], cwd=a, check=True)
    subprocess.run([os.path.join(ROOT, "damar_b200", "host", "simdb"), str(b), "Q"] + args + ["-X700", "-S1"], check=True,
                   stdout=subprocess.DEVNULL)
    assert open(a / ".Q.bps", "rb").read() == open(b / ".Q.bps", "rb").read()
    assert open(a / "Q.db").read() == open(b / "Q.db").read()
    x, y = dbio.read_block(str(a / "Q.2")), dbio.read_block(str(b / "Q.2"))
    assert np.array_equal(x.rlen, y.rlen) and np.array_equal(x.bases, y.bases) and np.array_equal(x.freq, y.freq)


def test_plan_matches_hpcdaligner(tmp_path):
    """Our job list is the reference planner's (HPCdaligner prints `daligner ... R.i R.1 .. R.i`)."""
    ref = os.path.join(ROOT, "oracle", "_ref", "HPCdaligner")
    if not os.path.exists(ref):
        pytest.skip("oracle/_ref not built")
    out = subprocess.run([ref, "-j4", os.path.join(TINY, "T")], capture_output=True, text=True).stdout
    cmds = [l.split() for l in out.splitlines() if l.startswith("daligner")]
    ours = plan.hpc_plan(2)
    got = {(c[-len(j.bs) - 1], tuple(c[-len(j.bs):])) for c, j in zip(sorted(cmds, key=lambda c: c[-1]), ours)}
    pairs_ref = set()
    for c in cmds:
        blocks = [t for t in c[1:] if not t.startswith("-")]
        a = int(blocks[0].split(".")[-1])
        pairs_ref |= {(a, int(b.split(".")[-1])) for b in blocks[1:]}
    pairs_ref = {(max(a, b), min(a, b)) for a, b in pairs_ref}
    assert pairs_ref == set(plan.block_pairs(2))
    assert got


@pytest.mark.parametrize("nblocks,world", [(1, 1), (4, 1), (4, 2), (4, 4), (4, 8), (21, 8), (7, 3)])
def test_sharding_covers_every_pair_once(nblocks, world):
    shards = plan.shard_pairs(nblocks, world)
    seen = [(j.a, b) for s in shards for j in s for b in j.bs]
    assert sorted(seen) == sorted(plan.block_pairs(nblocks))
    assert shards == plan.shard_pairs(nblocks, world)          # deterministic
    if nblocks * (nblocks + 1) // 2 >= 4 * world:
        loads = [sum(plan.pair_cost(j.a, b, {i: 1 for i in range(1, nblocks + 1)}) for j in s for b in j.bs) for s in shards]
        assert max(loads) <= 1.6 * (sum(loads) / world)
    owners = plan.block_owners(shards)
    for b, (owner, ranks) in owners.items():
        assert owner == ranks[0] and all(b in plan.blocks_needed(shards[r]) for r in ranks)


def _exchange_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    from damar_b200 import multi
    dist.init_process_group("gloo", rank=rank, world_size=world)
    shards = plan.shard_pairs(2, world)
    loads = []

    def loader(path, masks):
        loads.append(os.path.basename(path))
        return dbio.read_block(path, masks)
    got = multi.exchange_blocks(TINY, "T", shards, rank, world, dist, "cpu", ["rnd"], loader)
    q.put((rank, loads, {b: (blk.nreads, int(blk.bases.astype(np.int64).sum()), int(blk.mask_data.astype(np.int64).sum()),
                             blk.sizeof_db()) for b, blk in got.items()}, [(j.a, j.bs) for j in shards[rank]]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_block_exchange_two_ranks_gloo(world):
    """World size 2 (and 4: more ranks than units, one rank idles) on CPU: each block is read from disk by exactly one
    rank; a rank that needs a block it does not own receives an identical copy through the broadcast."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + 7 * world) % 2000
    ps = [ctx.Process(target=_exchange_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    truth = {b: dbio.read_block(os.path.join(TINY, f"T.{b}"), ["rnd"]) for b in (1, 2)}
    loaded = sorted(x for _, loads, _, _ in res for x in loads)
    assert loaded == ["T.1", "T.2"]                      # every block read from disk exactly once
    pairs = sorted((a, b) for _, _, _, jobs in res for a, bs in jobs for b in bs)
    assert pairs == [(1, 1), (2, 1), (2, 2)]
    for rank, _, blocks, jobs in res:
        need = {a for a, _ in jobs} | {b for _, bs in jobs for b in bs}
        assert set(blocks) == need
        for b, (n, s, ms, sz) in blocks.items():
            t = truth[b]
            assert (n, s, ms, sz) == (t.nreads, int(t.bases.astype(np.int64).sum()), int(t.mask_data.astype(np.int64).sum()), t.sizeof_db())


def test_las_merge_matches_reference_lamerge():
    """damar_las_merge (host) against the UNMODIFIED reference LAmerge on committed reference .las files: same bytes
    (tests/golden/lamerge.json holds size + sha256 of LAmerge's outputs, tests/golden/make_lamerge.py)."""
    import hashlib
    import json
    with open(os.path.join(GOLDEN, "lamerge.json")) as f:
        gold = json.load(f)
    for name, g in gold.items():
        ins = [open(os.path.join(GOLDEN, g["db"], "ref", fn), "rb").read() for fn in g["inputs"]]
        out = api.las_merge(ins)
        assert len(out) == g["size"], name
        assert hashlib.sha256(out).hexdigest() == g["sha256"], name
    assert gold["T.2"]["sha256"] == gold["T.2.rev"]["sha256"]
    a = open(os.path.join(GOLDEN, "tiny", "ref", "default", "T.1.T.1.las"), "rb").read()
    b = open(os.path.join(GOLDEN, "tiny", "ref", "k16s200", "T.1.T.1.las"), "rb").read()
    with pytest.raises(ValueError, match="trace spacings"):
        api.las_merge([a, b])
    with pytest.raises(ValueError, match="malformed"):
        api.las_merge([a[:-3]])


def test_marvel_track_loader(tmp_path):
    """track_load (lib/tracks.c:20-126): a MARVEL `.a2/.d2` track is preferred over the Dazzler `.anno/.data` pair and cut to
    the block.  The fixture is the tiny DB's random-interval track rewritten in MARVEL's format (dbio.write_marvel_track; the
    reference daligner reads such a file and reproduces the Dazzler-track case byte for byte -- checked when the writer was
    made); the oracle run with it must equal the reference's `rndmask` files."""
    import shutil
    import struct
    import oracle_ffi as O
    from cases import CASES, TINY, golden_files, oracle_kwargs, parse_flags
    for fn in os.listdir(TINY):
        if os.path.isfile(os.path.join(TINY, fn)):
            shutil.copy(os.path.join(TINY, fn), tmp_path / fn)
    with open(tmp_path / ".T.rnd.anno", "rb") as f:
        n, size = struct.unpack("<ii", f.read(8))
        anno = np.frombuffer(f.read(), dtype="<i8")
    data = np.fromfile(tmp_path / ".T.rnd.data", dtype="<i4")
    dbio.write_marvel_track(str(tmp_path / ".T"), "mrnd", anno, data)
    for name in ("T.1", "T.2"):
        d = dbio.read_block(str(tmp_path / name), ["rnd"])
        m = dbio.read_block(str(tmp_path / name), ["mrnd"])
        assert np.array_equal(d.mask_anno, m.mask_anno) and np.array_equal(d.mask_data, m.mask_data)
        assert m.mask_data.size > 100
    o = parse_flags(CASES["rndmask"]["flags"])
    blks = [dbio.read_block(str(tmp_path / nm), ["mrnd"]) for nm in ("T.1", "T.2")]
    for b in blks:
        b.track_names = ["rnd"]                       # sizeof_DB counts the track's name (irrelevant with -M0, kept equal anyway)
    out = O.daligner(blks[0], [blks[0], blks[1]], oracle_kwargs(o))
    for fn, ref in golden_files("rndmask").items():
        ok, why = dbio.las_equal(ref, out[fn])
        assert ok, f"{fn}: {why}"


def test_plan_iter_keeps_the_pool_fed_and_the_order(monkeypatch):
    """api.plan_iter (the plan runner behind daligner_plan and multi.run_sharded): jobs are collected in plan order, later jobs
    are submitted while at most DAMAR_LOOKAHEAD comparisons of the pending jobs are unfinished, and jobs whose process-wide
    options differ are never in flight together."""
    from concurrent.futures import Future
    from types import SimpleNamespace
    log, live = [], []

    class Blk(SimpleNamespace):
        pass

    def fake_submit(ablk, bblks, s, dcache=None):
        J = SimpleNamespace(name=ablk.name)
        J.cmps = [(b, False, None, {}, [Future(), Future()], {"files": {f"{ablk.name}.{b.name}.las": b"x"}}) for b in bblks]
        log.append(("submit", ablk.name, sum(1 for P in live for c in P.cmps for f in c[4] if not f.done())))
        live.append(J)
        return J

    def fake_collect(J, stats=None, as_bytes=True):
        for c in J.cmps:                       # the pool "finishes" a job's comparisons when the caller waits for them
            for f in c[4]:
                f.set_result(None)
        live.remove(J)
        log.append(("collect", J.name))
        return {k: v for c in J.cmps for k, v in c[5]["files"].items()}

    monkeypatch.setattr(api, "submit", fake_submit)
    monkeypatch.setattr(api, "collect", fake_collect)
    monkeypatch.setattr(api, "workers", lambda: 2)
    monkeypatch.setenv("DAMAR_LOOKAHEAD", "4")
    blk = lambda i, root="R": Blk(name=f"{root}.{i}", root=root, freq=(.25, .25, .25, .25))
    jobs = [(blk(a), [blk(b) for b in range(1, a + 1)]) for a in range(1, 6)]          # HPCdaligner's plan of 5 blocks
    files = list(api.plan_iter(jobs, api.Settings()))
    assert [k for k, _ in files] == [f"R.{a}.R.{b}.las" for a in range(1, 6) for b in range(1, a + 1)]
    assert [e[1] for e in log if e[0] == "collect"] == [f"R.{a}" for a in range(1, 6)]
    # a job is submitted only while fewer than 4 comparisons of MORE THAN ONE pending job are unfinished
    pend_at_submit = [e[2] for e in log if e[0] == "submit"]
    assert pend_at_submit[0] == 0 and pend_at_submit[1] == 2            # R.2 goes in while R.1 (one pair, two orientations) runs
    assert max(pend_at_submit) <= 2 * 4                                  # never more than the job being collected plus the lookahead
    # a second DB (other base frequencies): nothing of the first plan is in flight when its first job is submitted
    log.clear()
    other = Blk(name="S.1", root="S", freq=(.3, .2, .2, .3))
    list(api.plan_iter(jobs[:3] + [(other, [other])], api.Settings()))
    i = log.index(("submit", "S.1", 0))
    assert [e[1] for e in log[:i] if e[0] == "collect"] == ["R.1", "R.2", "R.3"]
