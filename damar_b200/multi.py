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


def run_sharded(dbdir: str, root: str, nblocks: int, settings, rank: int, world: int, dist=None, device: str = "cpu",
                outdir: Optional[str] = None, masks: Optional[List[str]] = None):
    """All-vs-all of a DB over `world` GPUs.  Returns (number of overlaps written by this rank, files)."""
    from . import api
    sizes = None
    shards = shard_pairs(nblocks, world, sizes)
    blocks = exchange_blocks(dbdir, root, shards, rank, world, dist, device, masks)
    cache: Dict[str, api.DeviceBlockHandle] = {}
    novl, files = 0, {}
    def finish(pending):
        nonlocal novl
        for name, im in api.collect(pending, as_bytes=False).items():
            novl += im.novl
            if outdir is not None:
                api.write_outputs(outdir, {name: im}, settings.run_id)
            files[name] = len(im)

    pend = None                                   # job i+1 is queued on the GPU before job i's records are sorted and written
    for job in shards[rank]:
        nxt = api.submit(blocks[job.a], [blocks[b] for b in job.bs], settings, dcache=cache)
        if pend is not None:
            finish(pend)
        pend = nxt
    if pend is not None:
        finish(pend)
    for h in cache.values():
        h.free()
    return novl, files
