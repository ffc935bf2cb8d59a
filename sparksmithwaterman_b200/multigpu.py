"""Reference-set sharding and the best-hit merge.

The pair set is refs x reads and pairs are independent (the reference itself partitions by
reference, Distribution.java:337), so each rank keeps a length-balanced shard of the
references resident in its HBM, aligns ALL reads against it, and only the per-read
best-hit records (score, global ref id, i, j) cross NVLink: one all_gather, then every rank
applies the same deterministic merge.  Works with torch.distributed on nccl (CUDA tensors)
and gloo (CPU tensors; used by the world_size-2 tests)."""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np


def shard_refs(lengths: Sequence[int], rank: int, world: int) -> List[int]:
    """Indices of the references rank `rank` owns: sort by descending length (stable), deal
    in snake order (0..w-1, w-1..0, ...) so no rank always gets the longest of a round;
    return ascending ids.  Every ref belongs to exactly one rank."""
    order = sorted(range(len(lengths)), key=lambda k: -int(lengths[k]))
    mine = []
    for pos, k in enumerate(order):
        rnd, slot = divmod(pos, world)
        if (slot if rnd % 2 == 0 else world - 1 - slot) == rank:
            mine.append(k)
    return sorted(mine)


def merge_best_hits(records: np.ndarray) -> np.ndarray:
    """records: [world, n_reads, 4] int (score, global ref id, i, j).  Winner per read =
    highest score, then lowest global ref id (the running-max-with-first-wins order of a
    single-GPU scan over refs in index order).  Returns [n_reads, 4]."""
    rec = np.asarray(records).astype(np.int64)
    key = rec[:, :, 0] * (1 << 32) - rec[:, :, 1]
    win = key.argmax(axis=0)
    return rec[win, np.arange(rec.shape[1])].astype(np.int32)


def localize(best_local: np.ndarray, my_ids: Sequence[int]) -> np.ndarray:
    """Shard-local ref indices -> global ref ids."""
    out = np.array(best_local, dtype=np.int32, copy=True)
    ids = np.asarray(my_ids, dtype=np.int32)
    ok = out[:, 1] >= 0
    out[ok, 1] = ids[out[ok, 1]]
    return out


def allgather_best_hits(best_global, group=None):
    """torch tensor [n_reads, 4] int32 (already global ref ids) -> merged [n_reads, 4] on every rank."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    out = [torch.empty_like(best_global) for _ in range(world)]
    dist.all_gather(out, best_global.contiguous(), group=group)
    allb = torch.stack(out).long()
    key = allb[:, :, 0] * (1 << 32) - allb[:, :, 1]
    win = key.argmax(dim=0)
    return allb[win, torch.arange(allb.shape[1], device=allb.device)].to(torch.int32)
