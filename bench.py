#!/usr/bin/env python
"""bench.py -- daligner overlaps/s on B200 (and the reference CPU daligner beside it).

Contract (see the task statement): `python bench.py --gpus N --steps K --warmup W` prints ONE JSON
line.  A *step* is one pass of the whole hot path over the workload: every job of the
HPCdaligner plan of the synthetic DB (all block pairs, both orientations), from k-mer tuples to
sorted `.las` images in host memory.

Workload (default, `config.workload`): BASELINE.json configs[1] -- simulator.c model, 4.6 Mbp genome
at 60x, 12% error, ~10 kbp reads, DBsplit -s70 -> 4 blocks, 10 block pairs, `daligner -k14 -w6 -h35
-e.70 -l1000 -s100 -M32`.  Data are synthetic (damar_b200/host/simdb restates db/simulator.c +
FA2db + DBsplit and writes a real Dazzler DB under a scratch directory; nothing is downloaded).

  value : overlaps/s with the blocks already resident in HBM when the timed region starts
  e2e   : overlaps/s through the public API with HOST buffers: every step uploads the unpacked
          blocks from host memory (H2D inside the timed region) and ends with the sorted .las
          images in host memory (D2H inside the timed region)
  roofline : the radix-sort scatter kernel (the dominant HBM-bound kernel), algorithmic bytes =
          32 B per element per pass on the reference's 16-byte records (SURVEY.md section 8d / DESIGN.md),
          duration measured with CUDA events around every launch, peak from MEASURED_PEAKS.json
  cpu_baseline : the unmodified reference daligner (oracle/_ref/daligner, `-j<P>`) on a bounded
          sample (one block pair) on this host's cores

Multi-GPU (N > 1, torchrun): block pairs are independent; every rank runs the same number of
units (its own copy of the plan on its own synthetic DB, seed = base + rank), no data-path
collective; value = total overlaps of all ranks / max-over-ranks device-timed duration ("weak").

`--impl reference` times the reference CPU daligner on the same config (bounded sample per step).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (simdb args, block size Mbp, root, daligner flags)
    "c2": dict(sim=["4.6", "-c60", "-r2", "-e.12", "-m10000", "-s2000", "-x4000"], split=70, root="E",
               desc="E. coli-scale 4.6 Mbp x60 (simulator.c model, 12% error), DBsplit -s70 -> 4 blocks, all 10 block pairs"),
    "c1": dict(sim=["5.0", "-c30", "-r1", "-e.12", "-m10000", "-s2000", "-x4000"], split=200, root="G",
               desc="5 Mbp x30 CLR-like, one block self-compare"),
    "small": dict(sim=["1.0", "-c20", "-r1", "-e.12"], split=10, root="S",
                  desc="1 Mbp x20, 2 blocks (smoke-sized)"),
    # BASELINE configs[2]: ONE DB, its block pairs sharded over the GPUs of the box (strong scaling), shared blocks broadcast
    # device to device over NCCL.  c3 = the configuration as named; c3q = the same plan shape (21 blocks, 231 pairs) at a
    # quarter of the bases, for development runs.
    "c3": dict(sim=["140", "-c30", "-r3", "-e.12", "-m10000", "-s2000", "-x4000"], split=200, root="D", sharded=True,
               desc="Drosophila-scale 140 Mbp x30 (simulator.c model, 12% error), DBsplit -s200 -> 21 blocks, all 231 block pairs sharded over the GPUs"),
    "c3q": dict(sim=["35", "-c30", "-r3", "-e.12", "-m10000", "-s2000", "-x4000"], split=50, root="Q", sharded=True,
                desc="35 Mbp x30 (quarter-scale configs[2]), DBsplit -s50 -> 21 blocks, all 231 block pairs sharded over the GPUs"),
    "c2s": dict(sim=["4.6", "-c60", "-r2", "-e.12", "-m10000", "-s2000", "-x4000"], split=70, root="E", sharded=True,
                desc="configs[1] data (4 blocks, 10 block pairs) run as ONE sharded all-vs-all"),
}
DAL_FLAGS = dict(k=14, w=6, h=35, e=0.70, l=1000, s=100, M=32)


def scratch_dir():
    for d in ("/dev/shm", "/tmp"):
        if os.path.isdir(d) and os.access(d, os.W_OK):
            return d
    return ROOT


def make_db(workload: str, seed_shift: int = 0):
    """Generate (once) the synthetic Dazzler DB for the workload; returns (dir, root, nblocks)."""
    w = WORKLOADS[workload]
    d = os.path.join(scratch_dir(), f"damar_bench_{workload}_{seed_shift}_{os.getuid()}")
    stub = os.path.join(d, w["root"] + ".db")
    if not os.path.exists(stub):
        os.makedirs(d, exist_ok=True)
        simdb = os.path.join(ROOT, "damar_b200", "host", "simdb")
        if not os.path.exists(simdb):
            subprocess.run(["make", "-s", "-C", ROOT, "host"], check=True)
        args = list(w["sim"])
        if seed_shift:
            args = [a if not a.startswith("-r") else f"-r{int(a[2:]) + 1000 * seed_shift}" for a in args]
        subprocess.run([simdb, d, w["root"]] + args + ["-X1000", f"-S{w['split']}"], check=True, stdout=subprocess.DEVNULL)
    from damar_b200 import dbio
    _, nblocks, _, _ = dbio.read_stub(stub)
    return d, w["root"], nblocks


def plan(nblocks: int):
    """HPCdaligner's job list (dalign/HPCdaligner.c:709-788, one job per A block):
    `daligner R.i R.1 ... R.i`."""
    return [(i, list(range(1, i + 1))) for i in range(1, nblocks + 1)]


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, gpu: int):
        self.gpu, self.proc, self.lines = gpu, None, []

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def roofline_lines(tm_alone, tm, dt, world, steps):
    """roofline = the dominant HBM-bound kernel, os_pass (one digit pass of the one-sweep radix sort), measured alone
    (tm_alone) and inside the timed steps (tm); stage4 = the kernel with the largest share of a step, wave_tasks, which is
    bound by integer issue slots, not by HBM or tensor cores."""
    peak, peak_src = peaks()
    sc_ms, sc_el, sc_n, sc_b = tm_alone["scatter_ms"], tm_alone["scatter_elems"], tm_alone["scatter_launches"], tm_alone["scatter_bytes"]
    alg_bytes = 32.0 * sc_el / max(sc_n, 1)                       # SURVEY 8d, per pass: 16 B read + 16 B write per record
    achieved = (32.0 * sc_el) / (sc_ms / 1000.0) / 1e9 if sc_ms > 0 else 0.0
    in_step = (32.0 * tm["scatter_elems"]) / (tm["scatter_ms"] / 1000.0) / 1e9 if tm["scatter_ms"] > 0 else 0.0
    whole = tm_alone["sort_contract_bytes"] / (tm_alone["sort_ms"] / 1000.0) / 1e9 if tm_alone["sort_ms"] > 0 else 0.0
    roofline = {"kernel": "os_pass (one-sweep radix sort: bulk-copy tile load, stable rank, look-back, scatter)", "bound": "hbm",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "measured": "CUDA events around every os_pass launch of the k-mer sorts of all blocks x orientations of the workload, "
                            "run one at a time after the timed steps (same data, same kernels)",
                "in_step": {"achieved": in_step, "frac": in_step / peak, "launches": tm["scatter_launches"],
                            "avg_launch_ms": tm["scatter_ms"] / max(tm["scatter_launches"], 1),
                            "share_of_step": (tm["scatter_ms"] / 1000.0) / dt if world == 1 else None,
                            "note": "event intervals inside the timed steps, several comparisons in flight: a launch also waits for "
                                    "SMs held by other comparisons' wave kernels"},
                # records are ONE 8-byte word here (16 B read + written per element and pass); dram bytes per element from
                # ncu --set full (profiles/r02_os_pass_ncu.txt): 8.1 read + 13.1 written (look-back words, partial sectors)
                "traffic": 21.2 * sc_el / max(sc_n, 1), "traffic_source": "ncu --set full dram__bytes_read+write per element x elements per launch",
                "moved_bytes_per_launch": sc_b / max(sc_n, 1),
                "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes, "launches": sc_n, "avg_launch_ms": sc_ms / max(sc_n, 1),
                "whole_sort": {"what": "histogram + scan + all passes of a k-mer sort against the contract bytes of SURVEY 8d "
                                       "(16-byte records, 8-bit digits: 16 n (1 + 2 ceil(bits / 8)))",
                               "achieved": whole, "frac": whole / peak, "elements": tm_alone["sort_elems"], "ms": tm_alone["sort_ms"]},
                # device time the kernel needs for the elements it sorted inside the timed steps (k-mer AND seed-hit sorts),
                # at the per-element rate measured alone, over the wall time of those steps: comparable with the kernel's
                # share of the ncu launch list (profiles/), where every launch also runs alone
                "share_of_step": ((sc_ms / max(sc_el, 1)) * tm["scatter_elems"] / 1000.0) / dt if world == 1 else None}
    ak = tm["align_kernel_ms"] / 1000.0
    stage4 = {"kernel": "wave_tasks (O(ND) wave extension, one thread per pass)", "bound": "integer issue (not HBM, no tensor cores)",
              "wave_cells_per_step": tm["cells_total"] // max(steps, 1), "local_alignments_per_step": tm["la_total"] // max(steps, 1),
              "wide_passes_per_step": tm["wide_total"] // max(steps, 1),
              "kernel_ms_per_step_summed_over_comparisons": 1000 * ak / max(steps, 1),
              "cells_per_s_per_launch": tm["cells_total"] / ak if ak > 0 else None,
              "cells_per_s_of_step": tm["cells_total"] / dt if dt > 0 else None,
              # the kernel running alone on one C2 block pair: committed ncu capture of this build (not measured in this run)
              "alone": {"cells": 803728653, "launch_ms": 28.02, "sms_full_ms": 16.0, "cells_per_s_while_full": 5.0e10,
                        "issue_active": 0.596, "lanes_active_of_32": 19.46, "registers": 154, "warps_per_sm": 12,
                        "source": "profiles/r02_ncu_wave_tasks.txt (ncu --set full), profiles/r02_wave_residency.txt (per-warp residency log)"},
              "note": "kernel_ms sums the event intervals of up to 6 concurrent comparisons, so cells_per_s_per_launch understates a launch "
                      "running alone; `alone` cites the committed capture of the same kernel"}
    return roofline, stage4


def ref_flags(threads: int):
    f = DAL_FLAGS
    return [f"-j{threads}", f"-k{f['k']}", f"-w{f['w']}", f"-h{f['h']}", f"-e{f['e']}", f"-l{f['l']}", f"-s{f['s']}", f"-M{f['M']}"]


def pow2_floor(n: int) -> int:
    p = 1
    while 2 * p <= n:
        p *= 2
    return p


def run_reference_job(dbdir: str, root: str, a: int, bs, threads: int, keep=None):
    """One reference daligner job (oracle/_ref/daligner, built from the unmodified reference).  keep: dict that receives
    {las file name: bytes} of what the job wrote (read after the clock stops)."""
    exe = os.path.join(ROOT, "oracle", "_ref", "daligner")
    if not os.path.exists(exe):
        raise RuntimeError("oracle/_ref/daligner missing (built by oracle/Makefile where /root/reference exists)")
    import shutil
    import struct
    for d in os.listdir(dbdir):
        if d.startswith("d001_"):
            shutil.rmtree(os.path.join(dbdir, d))
    t0 = time.perf_counter()
    subprocess.run([exe] + ref_flags(threads) + [f"{root}.{a}"] + [f"{root}.{b}" for b in bs], cwd=dbdir, check=True,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    dt = time.perf_counter() - t0
    novl = 0
    for d in os.listdir(dbdir):
        if d.startswith("d001_"):
            for fn in os.listdir(os.path.join(dbdir, d)):
                with open(os.path.join(dbdir, d, fn), "rb") as f:
                    if keep is not None:
                        keep[fn] = f.read()
                        novl += struct.unpack_from("<q", keep[fn])[0]
                    else:
                        novl += struct.unpack("<q", f.read(8))[0]
            shutil.rmtree(os.path.join(dbdir, d))
    return novl, dt


def main_sharded(args, rank, world, local_rank, cores, config, torch, dist, api, dbio):
    """ONE DB, its block pairs sharded over the ranks (plan.shard_pairs), every block read by one owner rank and broadcast
    PACKED, device to device, to the other ranks that need it (multi.exchange_blocks_device).  Strong scaling: the work is
    fixed, value = overlaps of the whole all-vs-all / max-over-ranks time."""
    from damar_b200 import multi
    from damar_b200 import plan as planmod
    device = f"cuda:{local_rank}"

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if rank == 0:
        make_db(args.workload)
    barrier()
    dbdir, root, nblocks = make_db(args.workload)
    s = api.Settings(kmer=DAL_FLAGS["k"], binshift=DAL_FLAGS["w"], hitmin=DAL_FLAGS["h"], ave_corr=DAL_FLAGS["e"],
                     minlen=DAL_FLAGS["l"], trace_spacing=DAL_FLAGS["s"], mem_limit_gb=DAL_FLAGS["M"], nthreads=4)
    shards = planmod.shard_pairs(nblocks, world)
    owners = planmod.block_owners(shards)
    host_blocks = {b: dbio.read_block(os.path.join(dbdir, f"{root}.{b}")) for b, (owner, _) in owners.items() if owner == rank}
    h2d = sum(int(b.bases.nbytes + b.boff.nbytes + b.rlen.nbytes) for b in host_blocks.values())
    for _ in range(max(1, args.warmup)):                           # whole passes incl. the exchange: NCCL groups, kernels, caches
        multi.run_sharded(dbdir, root, nblocks, s, rank, world, dist=dist, device=device, host_blocks=host_blocks)
    st = {}
    resident = multi.exchange_blocks_device(dbdir, root, shards, rank, world, dist, device, host_blocks=host_blocks, stats=st)
    sampler = ClockSampler(local_rank)
    api.reset_kernel_timers()
    l0 = api.timings()["kernel_launches"]
    barrier()
    sampler.start()
    t0 = time.perf_counter()
    novl = 0
    for _ in range(args.steps):
        n, files = multi.run_sharded(dbdir, root, nblocks, s, rank, world, dist=dist, device=device, resident=resident)
        novl += n
    barrier()
    dt = time.perf_counter() - t0
    tm = api.timings()
    launches = tm["kernel_launches"] - l0
    for h in resident[1].values():
        h.free()
    barrier()
    t1 = time.perf_counter()
    novl_e, d2h = 0, 0
    st_e = {}
    for _ in range(args.steps):
        n, files = multi.run_sharded(dbdir, root, nblocks, s, rank, world, dist=dist, device=device, host_blocks=host_blocks, stats=st_e)
        novl_e += n
        d2h = sum(files.values())
    barrier()
    dt_e = time.perf_counter() - t1
    clocks = sampler.stop()
    nvl = st.get("nvlink_bytes", 0)
    if world > 1:
        t = torch.tensor([dt, dt_e, float(novl), float(novl_e), float(launches), float(h2d), float(d2h), float(nvl)], dtype=torch.float64, device="cuda")
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        dt, dt_e = float(tmax[0]), float(tmax[1])
        novl, novl_e, launches = int(tsum[2]), int(tsum[3]), int(tsum[4])
        h2d, d2h, nvl = int(tsum[5]), int(tsum[6]), int(tsum[7])
    roofline, stage4 = roofline_lines(tm, tm, dt, world, args.steps)
    roofline["measured"] = "CUDA events around every os_pass launch inside the timed steps of rank 0 (several comparisons in flight)"
    config = dict(config, block_pairs=nblocks * (nblocks + 1) // 2, blocks=nblocks,
                  pairs_per_rank=[sum(len(j.bs) for j in sh) for sh in shards],
                  exchange="owner reads a block, packs it on its GPU and broadcasts the packed words device to device (NCCL) "
                           "to the ranks that also need it")
    line = {"metric": "daligner overlaps/s", "value": novl / dt, "unit": "overlaps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1000 * dt / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "int32", "data": "synthetic", "config": config,
            "e2e": {"value": novl_e / dt_e, "unit": "overlaps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": 1000 * dt_e / args.steps, "nvlink_bytes_per_step": nvl,
                    "exchange_ms_per_step_rank0": 1000 * st_e.get("exchange_s", 0.0) / max(args.steps, 1)},
            "all_vs_all": {"wall_ms": 1000 * dt_e / args.steps, "overlaps": novl_e // max(args.steps, 1), "n_gpus": world,
                           "what": "whole all-vs-all from host buffers on the owner ranks to sorted .las images in host memory"},
            "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "roofline_stage4": stage4,
            "overlaps_per_step": novl // max(args.steps, 1)}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:                                                     # bounded sample: one non-self block pair, byte-compared
            P = pow2_floor(cores)
            a, b = min(2, nblocks), 1
            ref_files = {}
            n_ref, t_ref = run_reference_job(dbdir, root, a, [b], P, keep=ref_files)
            blks = {i: (host_blocks.get(i) or dbio.read_block(os.path.join(dbdir, f"{root}.{i}"))) for i in {a, b}}
            got = api.daligner(blks[a], [blks[b]], s, as_bytes=False)
            bad = [fn for fn, ref in ref_files.items() if fn not in got or api.las_compare(ref, got[fn]) != -1]
            line["cpu_baseline"] = {"value": n_ref / t_ref, "unit": "overlaps/s", "cores": P, "kind": "reference",
                                    "sample": f"reference `daligner -j{P} {root}.{a} {root}.{b}` (1 of {nblocks * (nblocks + 1) // 2} block pairs, "
                                              f"{n_ref} overlaps in {t_ref:.1f} s) on {cores} host cores"}
            line["parity"] = {"checked": True, "files": len(ref_files), "bytes": sum(len(v) for v in ref_files.values()),
                              "overlaps": n_ref, "mismatches": bad[:8], "what": "the sampled block pair vs the reference daligner run on this host"}
        except Exception as e:  # noqa: BLE001
            line["cpu_baseline"] = {"value": None, "unit": "overlaps/s", "cores": 0, "kind": "reference", "sample": f"unavailable: {e}"}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    if line.get("parity", {}).get("mismatches"):
        sys.exit(3)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cores = os.cpu_count() or 1
    config = {"workload": args.workload, "desc": WORKLOADS[args.workload]["desc"],
              "daligner": "-k14 -w6 -h35 -e.70 -l1000 -s100 -M32", "l2": "inputs larger than L2 (every block's tuple list > 126 MB)"}

    # ---------------- reference arm: the reference's own CPU implementation ----------------
    if args.impl == "reference":
        if rank != 0:
            return
        dbdir, root, nblocks = make_db(args.workload)
        P = pow2_floor(cores)
        jobs = plan(nblocks)
        if WORKLOADS[args.workload].get("sharded") and nblocks > 4:
            # 231 block pairs of 200 Mbp blocks are hours of CPU: a step is ONE block pair, round-robin over a fixed sample
            jobs = [(1, [1]), (2, [1]), (nblocks, [nblocks - 1]), (nblocks // 2, [nblocks // 2])]
        # A step is ONE job of the HPCdaligner plan, taken round-robin (step i runs job i mod J), so K steps cover the whole
        # plan K/J times and the run stays within a few minutes.  value = overlaps of the whole plan / sum of the mean wall
        # time of every job = the reference's throughput on the SAME workload the GPU arm runs per step.
        run_reference_job(dbdir, root, jobs[0][0], jobs[0][1], P) if args.warmup > 0 else None
        tsum = [0.0] * len(jobs); cnt = [0] * len(jobs); nov = [0] * len(jobs)
        for i in range(args.steps):
            j = i % len(jobs)
            n, t = run_reference_job(dbdir, root, jobs[j][0], jobs[j][1], P)
            tsum[j] += t; cnt[j] += 1; nov[j] = n
        done = [j for j in range(len(jobs)) if cnt[j]]
        plan_s = sum(tsum[j] / cnt[j] for j in done)
        v = sum(nov[j] for j in done) / plan_s
        if WORKLOADS[args.workload].get("sharded") and nblocks > 4:
            config["reference_sample"] = "block pairs (1,1), (2,1), (n,n-1), (n/2,n/2) of the %d-pair plan, one per step" % (nblocks * (nblocks + 1) // 2)
        sample = (f"step i = job (i mod {len(jobs)}) of the {len(jobs)}-job plan (`daligner -j{P} {root}.a {root}.1 ... {root}.a`), "
                  f"{len(done)} of {len(jobs)} jobs sampled; whole plan = {nblocks * (nblocks + 1) // 2} block pairs, "
                  f"{sum(nov)} overlaps in {plan_s:.1f} s on {cores} host cores")
        print(json.dumps({"impl": "reference", "metric": "daligner overlaps/s", "value": v, "unit": "overlaps/s",
                          "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000 * sum(tsum) / max(args.steps, 1),
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32",
                          "data": "synthetic", "config": config, "plan_wall_s": plan_s,
                          "cpu_baseline": {"value": v, "unit": "overlaps/s", "cores": P, "kind": "reference", "sample": sample},
                          "e2e": {"value": v, "unit": "overlaps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    # ---------------- our arm ----------------
    import torch
    import torch.distributed as dist
    from damar_b200 import api, dbio

    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    torch.cuda.set_device(local_rank)
    api.init(local_rank)
    if WORKLOADS[args.workload].get("sharded"):
        return main_sharded(args, rank, world, local_rank, cores, config, torch, dist, api, dbio)

    dbdir, root, nblocks = make_db(args.workload, seed_shift=rank)
    blocks = {i: dbio.read_block(os.path.join(dbdir, f"{root}.{i}")) for i in range(1, nblocks + 1)}
    s = api.Settings(kmer=DAL_FLAGS["k"], binshift=DAL_FLAGS["w"], hitmin=DAL_FLAGS["h"], ave_corr=DAL_FLAGS["e"],
                     minlen=DAL_FLAGS["l"], trace_spacing=DAL_FLAGS["s"], mem_limit_gb=DAL_FLAGS["M"], nthreads=4)
    jobs = plan(nblocks)
    h2d = sum(int(b.bases.nbytes + b.boff.nbytes + b.rlen.nbytes) for b in blocks.values())

    last_files = {}

    def one_pass(resident, keep=False):
        """All jobs of the plan.  resident=None -> blocks are uploaded inside (e2e)."""
        novl, d2h = 0, 0
        cache = dict(resident) if resident is not None else {}
        for h in cache.values():
            h.drop_indexes()                     # a step always sorts its k-mers: only the packed blocks stay resident
        files = api.daligner_plan([(blocks[a], [blocks[b] for b in bs]) for a, bs in jobs], s, dcache=cache, as_bytes=False)
        for v in files.values():
            novl += v.novl
            d2h += len(v)
        if keep:
            last_files.clear(); last_files.update(files)
        if resident is None:
            for h in cache.values():
                h.free()
        return novl, d2h

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # warm-up (untimed): builds caches, loads kernels
    resident = {blocks[i].name: api.DeviceBlockHandle(blocks[i]) for i in blocks}
    for _ in range(args.warmup):
        one_pass(resident)

    sampler = ClockSampler(local_rank)
    # ---- device-resident timing ("value") ----
    api.load_library().damar_dump_profile() if os.environ.get("DAMAR_PROFILE") else None
    api.reset_kernel_timers()
    l0 = api.timings()["kernel_launches"]
    barrier()
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    t0 = time.perf_counter()
    novl = 0
    for _ in range(args.steps):
        n, d2h = one_pass(resident)
        novl += n
    ev1.record()
    barrier()
    dt = time.perf_counter() - t0
    dt_dev = ev0.elapsed_time(ev1) / 1000.0
    tm = api.timings()
    if os.environ.get("DAMAR_PROFILE"):
        print(f"[bench] host profile of {args.steps} resident steps:", file=sys.stderr)
        api.load_library().damar_dump_profile()
    launches = tm["kernel_launches"] - l0

    # ---- the dominant HBM kernel on its own: Sort_Kmers of every block and orientation of the workload, one at a
    # time.  Inside a step six comparisons are in flight, and the event interval around a sort launch then also
    # counts the time its CTAs wait for SMs held by another comparison's wave kernel; both figures are reported.
    api.reset_kernel_timers()
    barrier()
    for _ in range(args.steps):
        for h in resident.values():
            for comp in (0, 1):
                idx = api.sort_kmers(h, comp)
                if idx is not None:
                    idx.free()
    barrier()
    tm_alone = api.timings()
    for h in resident.values():
        h.free()

    # ---- end-to-end timing through the public API with host buffers ("e2e") ----
    barrier()
    t1 = time.perf_counter()
    novl_e = 0
    for i in range(args.steps):
        n, d2h = one_pass(None, keep=(i == args.steps - 1))
        novl_e += n
    barrier()
    dt_e = time.perf_counter() - t1
    clocks = sampler.stop()
    if os.environ.get("DAMAR_PROFILE"):
        print(f"[bench] host profile of {args.steps} end-to-end steps (after the sort-only region):", file=sys.stderr)
        api.load_library().damar_dump_profile()

    # ---- N > 1: one all-vs-all of ONE DB (rank 0's) with its block pairs sharded over the ranks (plan.shard_pairs) and the
    # blocks that several ranks need broadcast over NCCL (multi.exchange_blocks): the wall time BASELINE.json's metric asks for.
    # Strong scaling of a 10-pair plan, reported beside the weak-scaling `value`; DAMAR_BENCH_AVA=0 skips it.
    ava = None
    if world > 1 and os.environ.get("DAMAR_BENCH_AVA", "1") != "0":
        try:
            from damar_b200 import multi
            dbdir0, root0, nb0 = make_db(args.workload, seed_shift=0)
            multi.run_sharded(dbdir0, root0, nb0, s, rank, world, dist=dist, device=f"cuda:{local_rank}")      # warm-up (NCCL groups)
            barrier()
            t2 = time.perf_counter()
            n_ava, _ = multi.run_sharded(dbdir0, root0, nb0, s, rank, world, dist=dist, device=f"cuda:{local_rank}")
            barrier()
            dt_ava = time.perf_counter() - t2
            ta = torch.tensor([dt_ava, float(n_ava)], dtype=torch.float64, device="cuda")
            tm_ = ta.clone(); dist.all_reduce(tm_, op=dist.ReduceOp.MAX)
            ts_ = ta.clone(); dist.all_reduce(ts_, op=dist.ReduceOp.SUM)
            ava = {"wall_ms": 1000 * float(tm_[0]), "overlaps": int(ts_[1]), "block_pairs": nb0 * (nb0 + 1) // 2, "n_gpus": world,
                   "scaling": "strong", "overlaps_per_s": int(ts_[1]) / float(tm_[0]),
                   "what": "whole HPCdaligner plan of one DB from block files on disk to sorted .las images in host memory: "
                           "pairs sharded over the ranks, shared blocks broadcast over NCCL"}
        except Exception as e:  # noqa: BLE001
            ava = {"error": str(e)[:300]}

    if world > 1:
        t = torch.tensor([dt, dt_e, float(novl), float(novl_e), float(launches), float(h2d), float(d2h)], dtype=torch.float64, device="cuda")
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        dt, dt_e = float(tmax[0]), float(tmax[1])
        novl, novl_e, launches = int(tsum[2]), int(tsum[3]), int(tsum[4])
        h2d, d2h = int(tsum[5]), int(tsum[6])                      # whole job, like value
    value = novl / dt
    e2e = novl_e / dt_e

    roofline, stage4 = roofline_lines(tm_alone, tm, dt, world, args.steps)
    line = {"metric": "daligner overlaps/s", "value": value, "unit": "overlaps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1000 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int32", "data": "synthetic", "config": config,
            "e2e": {"value": e2e, "unit": "overlaps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": 1000 * dt_e / args.steps},
            "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "roofline_stage4": stage4,
            "overlaps_per_step": novl // max(args.steps, 1) // max(world, 1),
            "device_event_ms_per_step_rank0": 1000 * dt_dev / args.steps,
            "align_kernel_ms_per_step": tm["align_kernel_ms"] / max(args.steps, 1)}

    if ava is not None:
        line["all_vs_all"] = ava
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # The reference's own CPU daligner on the WHOLE plan of this workload, once (about 25 s on 16 cores), and every .las
        # file it writes byte-compared with the image the last timed end-to-end step of the GPU arm produced for the same
        # block pair (struct padding bytes 36..39 of a record excepted: the reference leaves them uninitialised).
        try:
            P = pow2_floor(cores)
            ref_files, n_ref, t_ref = {}, 0, 0.0
            for a, bs in jobs:
                n, t = run_reference_job(dbdir, root, a, bs, P, keep=ref_files)
                n_ref += n; t_ref += t
            line["cpu_baseline"] = {"value": n_ref / t_ref, "unit": "overlaps/s", "cores": P, "kind": "reference",
                                    "sample": f"reference `daligner -j{P}`, all {len(jobs)} jobs of the plan ({len(jobs) * (len(jobs) + 1) // 2} "
                                              f"block pairs, {n_ref} overlaps in {t_ref:.1f} s) on {cores} host cores"}
            bad, nbytes = [], 0
            for fn, ref in sorted(ref_files.items()):
                got = last_files.get(fn)
                if got is None:
                    bad.append(f"{fn}: missing"); continue
                where = api.las_compare(ref, got)
                nbytes += len(ref)
                if where != -1:
                    bad.append(f"{fn}: differs at byte {where}" if where >= 0 else f"{fn}: length {len(got)} != {len(ref)}")
            extra = sorted(set(last_files) - set(ref_files))
            bad += [f"{fn}: not written by the reference" for fn in extra]
            line["parity"] = {"checked": True, "files": len(ref_files), "bytes": nbytes, "overlaps": n_ref, "mismatches": bad[:8],
                              "what": "every .las of the last end-to-end step vs the reference daligner run on this host"}
        except Exception as e:  # noqa: BLE001
            line["cpu_baseline"] = {"value": None, "unit": "overlaps/s", "cores": 0, "kind": "reference", "sample": f"unavailable: {e}"}
            line["parity"] = {"checked": False, "why": str(e)[:200]}
    last_files.clear()
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    if line.get("parity", {}).get("mismatches"):
        sys.exit(3)                                        # a fast run whose output differs from the reference's is not a result


if __name__ == "__main__":
    main()
