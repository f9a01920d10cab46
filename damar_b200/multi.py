"""Multi-GPU execution of a sharded plan: one process per GPU (torchrun), block pairs sharded with
`plan.shard_pairs`, read blocks broadcast between ranks only when several ranks need the same block
(torch.distributed: NCCL over NVLink on GPUs, gloo on CPU for the host-logic tests).  There is no
data-path collective: results are per-pair .las files written by the owning rank.
"""
from __future__ import annotations

import os
from typing import Callable, Dict, List, Optional, Sequence

import numpy as np

from . import dbio
from .plan import Job, block_owners, shard_pairs

_FIELDS = ("rlen", "boff", "bases", "freq")


def _pack_meta(blk: dbio.Block) -> np.ndarray:
    has_mask = blk.mask_anno is not None
    return np.asarray([blk.part, blk.ufirst, blk.nreads, blk.maxlen, blk.totlen, blk.ureads, blk.bases.size,
                       int(has_mask), blk.mask_data.size if has_mask else 0], dtype=np.int64)


_GROUPS: Dict[tuple, object] = {}


def _group_for(dist, ranks, world):
    """Process group of exactly `ranks`, created once per process (an NCCL communicator takes seconds to set up);
    None = the default group when every rank takes part."""
    key = tuple(ranks)
    if len(key) == world:
        return None
    if key not in _GROUPS:
        _GROUPS[key] = dist.new_group(ranks=list(key))
    return _GROUPS[key]


def exchange_blocks(dbdir: str, root: str, shards: Sequence[Sequence[Job]], rank: int, world: int,
                    dist=None, device: str = "cpu", masks: Optional[List[str]] = None,
                    loader: Callable[[str, Optional[List[str]]], dbio.Block] = dbio.read_block) -> Dict[int, dbio.Block]:
    """Return the host blocks this rank needs.  Each block is read from disk by its owner only;
    other ranks that need it receive it through a broadcast in a sub-group of exactly those ranks."""
    import torch
    owners = block_owners(shards)
    mine: Dict[int, dbio.Block] = {}
    for b, (owner, ranks) in owners.items():
        blk = None
        if rank == owner:
            blk = loader(os.path.join(dbdir, f"{root}.{b}"), masks)
        if len(ranks) > 1 and world > 1:
            group = _group_for(dist, ranks, world)          # every rank must create every group
            if rank in ranks:
                meta = torch.from_numpy(_pack_meta(blk) if blk is not None else np.zeros(9, np.int64)).to(device)
                dist.broadcast(meta, src=owner, group=group)
                m = meta.cpu().numpy()
                arrs = {}
                spec = [("rlen", np.int32, int(m[2])), ("boff", np.int64, int(m[2]) + 1), ("bases", np.uint8, int(m[6])),
                        ("freq", np.float32, 4)]
                if m[7]:
                    spec += [("mask_anno", np.int64, int(m[2]) + 1), ("mask_data", np.int32, max(int(m[8]), 1))]
                for name, dt, n in spec:
                    if blk is not None:
                        src = getattr(blk, name)
                        if name == "mask_data" and src.size == 0:
                            src = np.zeros(1, np.int32)
                        t = torch.from_numpy(np.ascontiguousarray(src, dtype=dt)).to(device)
                    else:
                        t = torch.empty(n, dtype=getattr(torch, np.dtype(dt).name), device=device)
                    dist.broadcast(t, src=owner, group=group)
                    arrs[name] = t.cpu().numpy()
                if blk is None:
                    blk = dbio.Block(name=f"{root}.{b}", root=root, path=os.path.join(dbdir, "." + root), part=int(m[0]),
                                     ufirst=int(m[1]), nreads=int(m[2]), maxlen=int(m[3]), totlen=int(m[4]), ureads=int(m[5]),
                                     freq=arrs["freq"], rlen=arrs["rlen"], boff=arrs["boff"], bases=arrs["bases"],
                                     mask_anno=arrs.get("mask_anno"),
                                     mask_data=(arrs["mask_data"][: int(m[8])] if m[7] else None),
                                     track_names=list(masks or []))
        if rank in ranks:
            mine[b] = blk
    return mine


class _DevArray:
    """A device buffer of the library seen as a CUDA array (zero copy), so torch.distributed can broadcast it."""

    def __init__(self, ptr: int, n: int, typestr: str):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


def exchange_blocks_device(dbdir: str, root: str, shards: Sequence[Sequence[Job]], rank: int, world: int, dist, device: str,
                           masks: Optional[List[str]] = None,
                           loader: Callable[[str, Optional[List[str]]], dbio.Block] = dbio.read_block,
                           host_blocks: Optional[Dict[int, dbio.Block]] = None, stats: Optional[dict] = None):
    """GPU path of exchange_blocks: the owner reads a block (or takes it from `host_blocks`), ships it to its HBM where it is
    packed to 2 bit/base, and broadcasts the PACKED words + read offsets / lengths (+ mask) device to device over NCCL, in a
    group of exactly the ranks that need the block; receivers adopt the buffers (damar_block_from_view).  Nothing unpacked
    crosses NVLink and nothing bounces through a host.  Returns ({block number: metadata Block}, {block name: handle})."""
    import torch
    from . import api
    owners = block_owners(shards)
    meta: Dict[int, dbio.Block] = {}
    handles: Dict[str, api.DeviceBlockHandle] = {}
    sent = 0
    for b, (owner, ranks) in owners.items():
        name = f"{root}.{b}"
        blk, h = None, None
        if rank == owner:
            blk = host_blocks[b] if host_blocks is not None else loader(os.path.join(dbdir, name), masks)
            h = api.DeviceBlockHandle(blk)
        if len(ranks) > 1 and world > 1:
            # a block most ranks need goes over the world group (the others drop it): a communicator per subset of ranks
            # costs more to set up than the 50 MB cost to send
            part = ranks if len(ranks) <= world // 2 else list(range(world))
            group = _group_for(dist, part, world)            # every rank must create every group
            if rank in part:
                if rank == owner:
                    v = h.view()
                    m = torch.tensor([v.nreads, v.maxlen, v.ufirst, v.part, v.totlen, v.nbases, v.sizeof_db, v.npacked_words,
                                      v.mask_len, int(bool(v.mask_anno)), v.flags, blk.ureads], dtype=torch.int64, device=device)
                    fq = torch.tensor([float(x) for x in v.freq], dtype=torch.float64, device=device)
                else:
                    m = torch.zeros(12, dtype=torch.int64, device=device)
                    fq = torch.zeros(4, dtype=torch.float64, device=device)
                dist.broadcast(m, src=owner, group=group)
                dist.broadcast(fq, src=owner, group=group)
                mm = [int(x) for x in m.cpu()]
                nreads, npw, mask_len, has_mask = mm[0], mm[7], mm[8], mm[9]
                spec = [("packed", torch.int32, "<i4", npw + 4), ("boff", torch.int32, "<i4", nreads + 1), ("rlen", torch.int32, "<i4", nreads)]
                if has_mask:
                    spec += [("mask_anno", torch.int64, "<i8", nreads + 1), ("mask_data", torch.int32, "<i4", max(mask_len, 1))]
                bufs = {}
                for nm, dt, ts, n in spec:
                    if rank == owner:
                        ptr = getattr(v, nm)
                        t = torch.as_tensor(_DevArray(ptr, n, ts), device=device) if ptr else torch.zeros(n, dtype=dt, device=device)
                    else:
                        t = torch.empty(n, dtype=dt, device=device)
                    dist.broadcast(t, src=owner, group=group)
                    sent += t.numel() * t.element_size()
                    bufs[nm] = t
                torch.cuda.synchronize()
                if rank != owner and rank in ranks:
                    nv = api.BlockView()
                    nv.nreads, nv.maxlen, nv.ufirst, nv.part, nv.totlen, nv.nbases, nv.sizeof_db, nv.npacked_words = mm[:8]
                    nv.mask_len, nv.flags = mask_len, mm[10]
                    fqh = fq.cpu()
                    for i in range(4):
                        nv.freq[i] = float(fqh[i])
                    nv.packed, nv.boff, nv.rlen = bufs["packed"].data_ptr(), bufs["boff"].data_ptr(), bufs["rlen"].data_ptr()
                    if has_mask:
                        nv.mask_anno, nv.mask_data = bufs["mask_anno"].data_ptr(), bufs["mask_data"].data_ptr()
                    rl = bufs["rlen"].cpu().numpy().astype(np.int32)
                    bo = bufs["boff"].cpu().numpy().astype(np.uint32).astype(np.int64)
                    blk = dbio.Block(name=name, root=root, path=os.path.join(dbdir, "." + root), part=mm[3], ufirst=mm[2],
                                     nreads=nreads, maxlen=mm[1], totlen=mm[4], ureads=mm[11],
                                     freq=np.asarray([float(x) for x in fqh], np.float32), rlen=rl, boff=bo, bases=None,
                                     track_names=list(masks or []))
                    h = api.DeviceBlockHandle.from_view(nv, blk)
                del bufs
        if rank in ranks:
            meta[b] = blk
            handles[name] = h
    if stats is not None:
        stats["nvlink_bytes"] = stats.get("nvlink_bytes", 0) + sent
    return meta, handles


def run_sharded(dbdir: str, root: str, nblocks: int, settings, rank: int, world: int, dist=None, device: str = "cpu",
                outdir: Optional[str] = None, masks: Optional[List[str]] = None, resident=None, stats: Optional[dict] = None,
                host_blocks: Optional[Dict[int, dbio.Block]] = None):
    """All-vs-all of a DB over `world` GPUs.  Returns (number of overlaps written by this rank, files).
    resident = (metadata blocks, handles) of an earlier exchange keeps the blocks in HBM across calls."""
    from . import api
    sizes = None
    shards = shard_pairs(nblocks, world, sizes)
    own_cache = resident is None
    if resident is not None:
        blocks, cache = resident
        for h in cache.values():
            h.drop_indexes()
    elif str(device).startswith("cuda"):
        import time
        t0 = time.perf_counter()
        blocks, cache = exchange_blocks_device(dbdir, root, shards, rank, world, dist, device, masks, host_blocks=host_blocks, stats=stats)
        if stats is not None:
            stats["exchange_s"] = stats.get("exchange_s", 0.0) + time.perf_counter() - t0
    else:
        blocks = exchange_blocks(dbdir, root, shards, rank, world, dist, device, masks)
        cache = {}
    novl, files = 0, {}
    # the rank's jobs through the library's pipelined plan runner (later jobs are queued while earlier ones finish)
    for name, im in api.plan_iter(((blocks[job.a], [blocks[b] for b in job.bs]) for job in shards[rank]), settings,
                                  dcache=cache, as_bytes=False):
        novl += im.novl
        if outdir is not None:
            api.write_outputs(outdir, {name: im}, settings.run_id)
        files[name] = len(im)
    if own_cache:
        for h in cache.values():
            h.free()
    return novl, files
