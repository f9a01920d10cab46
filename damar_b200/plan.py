"""HPCdaligner-style job planning and its sharding over GPUs.

The reference's planner (dalign/HPCdaligner.c:572-812) prints one `daligner` command per A block
(or per --dal group) comparing block i against blocks 1..i, then one LAmerge per block.  Block pairs
are independent units of work: {X,Y}, X >= Y, both orientations, emitting X.Y.las and
Y.X.las.  Here the same pair list is built and partitioned over ranks; a rank keeps the k-mer index
of its current A block and walks that block's B partners, exactly like one daligner process does.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Sequence, Set, Tuple


@dataclass(frozen=True)
class Job:
    """One `daligner A B1 B2 ...` invocation."""
    a: int
    bs: Tuple[int, ...]

    def command(self, root: str, flags: Sequence[str] = ()) -> str:
        return " ".join(["daligner", *flags, f"{root}.{self.a}", *[f"{root}.{b}" for b in self.bs]])


def block_pairs(nblocks: int, first: int = 1, last: int | None = None) -> List[Tuple[int, int]]:
    """All units (a, b), a >= b, for A blocks first..last (HPCdaligner.c:631-788)."""
    last = nblocks if last is None else last
    return [(a, b) for a in range(first, last + 1) for b in range(1, a + 1)]


def hpc_plan(nblocks: int, dal: int = 0) -> List[Job]:
    """Jobs in the reference's order: A = 1..n, B partners 1..A in groups of `dal` (0 = all in one job)."""
    jobs = []
    for a in range(1, nblocks + 1):
        bs = list(range(1, a + 1))
        if dal <= 0:
            jobs.append(Job(a, tuple(bs)))
        else:
            for i in range(0, len(bs), dal):
                jobs.append(Job(a, tuple(bs[i:i + dal])))
    return jobs


def pair_cost(a: int, b: int, sizes: Dict[int, int]) -> float:
    """Relative cost of a unit: seed hits scale with |A|*|B|; a self pair only uses the lower triangle."""
    c = float(sizes[a]) * float(sizes[b])
    return 0.5 * c if a == b else c


def shard_pairs(nblocks: int, world: int, sizes: Dict[int, int] | None = None) -> List[List[Job]]:
    """Partition all block pairs over `world` ranks.

    Units are grouped by A block so a rank reuses the A index (the reference's --dal idea), and A blocks
    are dealt out longest-processing-time-first.  When there are fewer A blocks than ranks would need for
    balance, an A block's partner list is split between ranks.  Deterministic; every pair appears once.
    """
    sizes = sizes or {i: 1 for i in range(1, nblocks + 1)}
    units = block_pairs(nblocks)
    total = sum(pair_cost(a, b, sizes) for a, b in units)
    # Pieces of a quarter of a rank's share: small enough for the longest-first deal to balance the ranks to a few percent
    # (21 blocks on 8 ranks: 0.90 -> 0.99 of ideal), large enough to keep a rank on one A block for a while.  A rank keeps
    # the k-mer indexes of all its blocks for the whole plan (api.DeviceBlockHandle.index), so a piece costs no extra sort.
    target = total / world / 4
    # split rows (A blocks) into chunks no bigger than that
    chunks: List[Tuple[float, int, List[int]]] = []
    for a in range(1, nblocks + 1):
        cur: List[int] = []
        acc = 0.0
        for b in range(1, a + 1):
            c = pair_cost(a, b, sizes)
            if cur and acc + c > target * 1.0001:
                chunks.append((acc, a, cur))
                cur, acc = [], 0.0
            cur.append(b)
            acc += c
        if cur:
            chunks.append((acc, a, cur))
    chunks.sort(key=lambda t: (-t[0], t[1], t[2][0]))
    load = [0.0] * world
    out: List[List[Job]] = [[] for _ in range(world)]
    for cost, a, bs in chunks:
        r = min(range(world), key=lambda i: (load[i], i))
        out[r].append(Job(a, tuple(bs)))
        load[r] += cost
    for r in range(world):
        out[r].sort(key=lambda j: (j.a, j.bs))
    return out


def blocks_needed(jobs: Sequence[Job]) -> Set[int]:
    need: Set[int] = set()
    for j in jobs:
        need.add(j.a)
        need.update(j.bs)
    return need


def block_owners(shards: Sequence[Sequence[Job]]) -> Dict[int, Tuple[int, List[int]]]:
    """block -> (owner rank that loads it from disk, sorted ranks that need it).  A block needed by
    several ranks is loaded once and broadcast (NCCL over NVLink); otherwise no communication."""
    need: Dict[int, List[int]] = {}
    for r, jobs in enumerate(shards):
        for b in blocks_needed(jobs):
            need.setdefault(b, []).append(r)
    return {b: (min(rs), sorted(rs)) for b, rs in sorted(need.items())}
