#!/usr/bin/env python
"""Reduced-size versions of BASELINE.json configs[3] (Nanopore-like long reads, -e.65) and configs[4] (HiFi-like, -k20 -e.95
-mdust) on the GPU box: ours vs the UNMODIFIED reference daligner (oracle/_ref/daligner), every .las compared bit for bit
(padding masked).  Test infrastructure (executes oracle/_ref).  usage: python tools/gpu_check_configs.py [c4s] [c5s]"""
import os
import shutil
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from damar_b200 import api, dbio  # noqa: E402

CONFIGS = {
    # configs[2]-like block size: two DBsplit -s200 blocks (2 x 200 Mbp, the production block size) of a 20 Mbp genome at 20x,
    # the non-self pair only: 200 M k-mers per list, chance 14-mer hits dominate
    "c3s": dict(sim=["20.0", "-c20", "-r3", "-e.12", "-m10000", "-s2000", "-x4000"], split=200, root="D",
                flags=["-k14", "-w6", "-h35", "-e.70", "-l1000", "-s100", "-M32"], masks=[], jobs=[(2, [1])]),
    # Nanopore-like: N50 ~ 30 kbp, 15 % error; long reads -> long passes, wide fronts (64-ring and strided fallbacks)
    "c4s": dict(sim=["6.0", "-c25", "-r4", "-e.15", "-m22000", "-s18000", "-x4000"], split=200, root="N",
                flags=["-k14", "-w6", "-h35", "-e.65", "-l1000", "-s100", "-M32"], masks=[]),
    # HiFi-like: 99.5 % identity, -k20 (k % 4 == 0, 5 sort passes), dust mask, -e.95: seed-heavy sort / merge
    "c5s": dict(sim=["8.0", "-c25", "-r5", "-e.005", "-m15000", "-s2000", "-x4000"], split=250, root="H",
                flags=["-k20", "-w6", "-h35", "-e.95", "-l1000", "-s100", "-M32", "-mdust"], masks=["dust"]),
}


def settings(flags):
    o = dict(k=14, w=6, h=35, e=.70, l=1000, s=100, M=32, t=0)
    sw = set()
    for f in flags:
        if f[1] in "kwhlsMt":
            o[f[1]] = int(f[2:])
        elif f[1] == "e":
            o["e"] = float(f[2:])
        elif f[1] in "IATOb":
            sw.add(f[1])
    return api.Settings(kmer=o["k"], binshift=o["w"], hitmin=o["h"], ave_corr=o["e"], minlen=o["l"], trace_spacing=o["s"],
                        mem_limit_gb=o["M"], nthreads=4, suppress=o["t"], identity="I" in sw, symmetric="A" not in sw,
                        no_trace="T" in sw, only_identity="O" in sw, biased="b" in sw)


# the option space at 2 x 10 Mbp (bench workload "small"): job `daligner [flags] S.2 S.1 S.2` against the reference binary
FLAG_SETS = [[], ["-t12"], ["-I"], ["-A"], ["-T"], ["-s200"], ["-s50"], ["-k16", "-h50"], ["-k12", "-w5", "-h28"], ["-e.85"],
             ["-l2500"], ["-b"], ["-M1"], ["-k20", "-h60", "-e.8"], ["-w7", "-h40"]]


def flag_sweep():
    dbdir, root, nblocks = bench.make_db("small")
    blocks = {i: dbio.read_block(os.path.join(dbdir, f"{root}.{i}")) for i in (1, 2)}
    allok = True
    for fl in FLAG_SETS:
        flags = ["-M32"] + fl if not any(f.startswith("-M") for f in fl) else list(fl)
        s = settings(flags)
        got = api.daligner(blocks[2], [blocks[1], blocks[2]], s)
        for d in os.listdir(dbdir):
            if d.startswith("d001_"):
                shutil.rmtree(os.path.join(dbdir, d))
        subprocess.run([os.path.join(ROOT, "oracle", "_ref", "daligner"), "-j4"] + flags + [f"{root}.2", f"{root}.1", f"{root}.2"],
                       cwd=dbdir, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        ok_all, novl = True, 0
        for fn, v in got.items():
            part = int(fn.split(".")[1])
            ref = open(os.path.join(dbdir, f"d001_{part:05d}", fn), "rb").read()
            ok, why = dbio.las_equal(ref, v)
            ok_all &= ok
            novl += int.from_bytes(v[:8], "little")
            if not ok:
                print(f"    {fn}: DIFF {why}", flush=True)
        allok &= ok_all
        print(f"  flags {' '.join(flags):28s} {len(got)} files, {novl:7d} overlaps: {'OK' if ok_all else 'DIFF'}", flush=True)
    return allok


def main():
    names = sys.argv[1:] or list(CONFIGS)
    api.init()
    allok = True
    if "flags" in names:
        names.remove("flags")
        allok &= flag_sweep()
    for name in names:
        cfg = CONFIGS[name]
        bench.WORKLOADS[name] = dict(sim=cfg["sim"], split=cfg["split"], root=cfg["root"], desc=name)
        dbdir, root, nblocks = bench.make_db(name)
        if cfg["masks"] and not os.path.exists(os.path.join(dbdir, f".{root}.dust.anno")):
            subprocess.run([os.path.join(ROOT, "oracle", "_ref", "DBdust"), root], cwd=dbdir, check=True, stdout=subprocess.DEVNULL)
        blocks = {i: dbio.read_block(os.path.join(dbdir, f"{root}.{i}"), cfg["masks"] or None) for i in range(1, nblocks + 1)}
        s = settings(cfg["flags"])
        P = bench.pow2_floor(os.cpu_count())
        print(f"{name}: {nblocks} block(s), {sum(b.totlen for b in blocks.values()) / 1e6:.0f} Mbp, flags {' '.join(cfg['flags'])}", flush=True)
        for a, bs in cfg.get("jobs") or bench.plan(nblocks):
            api.daligner(blocks[a], [blocks[b] for b in bs], s, as_bytes=False)          # warm-up
            stats = []
            t0 = time.time()
            got = api.daligner(blocks[a], [blocks[b] for b in bs], s, stats=stats)
            t_gpu = time.time() - t0
            for st in stats:
                for comp in (0, 1):
                    t = st[f"t{comp}"]
                    print(f"  job {a} comp{comp}: kmers={t['kmers']} hits={st[f'nhits{comp}']} la={st[f'nfilt{comp}']} ovl={st[f'novl{comp}']} multi={st[f'multi{comp}']}"
                          f" | ksort {t['ksort_ms']:.1f} count {t['count_ms']:.1f} merge {t['merge_ms']:.1f} hsort {t['hsort_ms']:.1f} pairs {t['pairs_ms']:.1f}"
                          f" align {t['align_ms']:.1f} ms redo={t['redo_pairs']}", flush=True)
            for d in os.listdir(dbdir):
                if d.startswith("d001_"):
                    shutil.rmtree(os.path.join(dbdir, d))
            t0 = time.time()
            subprocess.run([os.path.join(ROOT, "oracle", "_ref", "daligner"), f"-j{P}"] + cfg["flags"] + [f"{root}.{a}"] + [f"{root}.{b}" for b in bs],
                           cwd=dbdir, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            t_ref = time.time() - t0
            novl = 0
            for fn, v in got.items():
                part = int(fn.split(".")[1])
                ref = open(os.path.join(dbdir, f"d001_{part:05d}", fn), "rb").read()
                ok, why = dbio.las_equal(ref, v)
                allok &= ok
                novl += int.from_bytes(v[:8], "little")
                print(f"  {fn}: {len(v)} bytes {'OK' if ok else 'DIFF ' + why}", flush=True)
            print(f"  job {a} vs {bs}: ours {t_gpu:.2f} s, reference -j{P} {t_ref:.2f} s, {novl} overlaps", flush=True)
    print("PARITY", "PASS" if allok else "FAIL")
    sys.exit(0 if allok else 1)


if __name__ == "__main__":
    main()
