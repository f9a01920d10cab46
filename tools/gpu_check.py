#!/usr/bin/env python
"""Full-size parity + timing check on the GPU box: ours vs the UNMODIFIED reference daligner
(oracle/_ref/daligner) on a bench workload, every .las compared bit for bit (padding masked).
Test infrastructure (executes oracle/_ref).  usage: python tools/gpu_check.py <workload> [jobs...]"""
import json
import os
import shutil
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from damar_b200 import api, dbio  # noqa: E402


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "small"
    only = [int(x) for x in sys.argv[2:]]
    dbdir, root, nblocks = bench.make_db(wl)
    api.init()
    f = bench.DAL_FLAGS
    s = api.Settings(kmer=f["k"], binshift=f["w"], hitmin=f["h"], ave_corr=f["e"], minlen=f["l"], trace_spacing=f["s"],
                     mem_limit_gb=f["M"], nthreads=4)
    blocks = {}
    t0 = time.time()
    for i in range(1, nblocks + 1):
        blocks[i] = dbio.read_block(os.path.join(dbdir, f"{root}.{i}"))
    print(f"loaded {nblocks} blocks in {time.time() - t0:.1f}s", flush=True)
    cores = os.cpu_count()
    P = bench.pow2_floor(cores)
    allok = True
    for a, bs in bench.plan(nblocks):
        if only and a not in only:
            continue
        api.daligner(blocks[a], [blocks[b] for b in bs], s, as_bytes=False)     # warm-up (allocator, kernels)
        api.load_library().damar_dump_profile()
        print("  -- second run --", flush=True)
        stats = []
        t0 = time.time()
        got = api.daligner(blocks[a], [blocks[b] for b in bs], s, stats=stats)
        t_gpu = time.time() - t0
        api.load_library().damar_dump_profile()
        for st in stats:
            for comp in (0, 1):
                t = st[f"t{comp}"]
                print(f"  job {a} comp{comp}: hits={st[f'nhits{comp}']} la={st[f'nfilt{comp}']} ovl={st[f'novl{comp}']} multi={st[f'multi{comp}']} "
                      f"| extract {t['extract_ms']:.1f} ksort {t['ksort_ms']:.1f} count {t['count_ms']:.1f} merge {t['merge_ms']:.1f} "
                      f"hsort {t['hsort_ms']:.1f} pairs {t['pairs_ms']:.1f} align {t['align_ms']:.1f} dl {t['download_ms']:.1f} ms; cells={t['wave_cells']} pairs={t['pairs']} redo={t['redo_pairs']} fastk={t['align_kernel_ms']:.1f} redok={t['redo_kernel_ms']:.1f} (cumulative)", flush=True)
        for d in os.listdir(dbdir):
            if d.startswith("d001_"):
                shutil.rmtree(os.path.join(dbdir, d))
        t0 = time.time()
        subprocess.run([os.path.join(ROOT, "oracle", "_ref", "daligner")] + bench.ref_flags(P) + [f"{root}.{a}"] + [f"{root}.{b}" for b in bs],
                       cwd=dbdir, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL,
                       env=dict(os.environ, OMP_NUM_THREADS=str(P)))
        t_ref = time.time() - t0
        novl = 0
        for fn, v in got.items():
            part = int(fn.split(".")[1])
            ref = open(os.path.join(dbdir, f"d001_{part:05d}", fn), "rb").read()
            ok, why = dbio.las_equal(ref, v)
            allok &= ok
            novl += int.from_bytes(v[:8], "little")
            print(f"  {fn}: {len(v)} bytes {'OK' if ok else 'DIFF ' + why}", flush=True)
        print(f"job {a} vs {bs}: ours {t_gpu:.2f}s  reference(-j{P}) {t_ref:.2f}s  overlaps {novl}", flush=True)
    print("PARITY", "PASS" if allok else "FAIL")
    sys.exit(0 if allok else 1)


if __name__ == "__main__":
    main()
