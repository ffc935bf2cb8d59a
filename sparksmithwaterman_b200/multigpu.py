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
    # a record without a reference (ref < 0: a rank whose shard is empty) loses against every real one
    key = rec[:, :, 0] * (1 << 32) - np.where(rec[:, :, 1] < 0, 1 << 31, rec[:, :, 1])
    key = np.where(rec[:, :, 1] < 0, -(1 << 62), key)
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
    key = torch.where(allb[:, :, 1] < 0, torch.full_like(key, -(1 << 62)), key)      # empty shard: never the winner
    win = key.argmax(dim=0)
    return allb[win, torch.arange(allb.shape[1], device=allb.device)].to(torch.int32)


# ---------------------------------------------------------------------------------------------------
# The same sharding + merge INSIDE the C ABI (include/swb200.h, "multi-GPU"): what a JVM host binds.
import ctypes as _C

from . import _ffi
from .engine import AlignResult, DEFAULT_SCORES, concat, _i64p
from ._ffi import SWB_F_NO_FETCH, SWB_F_SCORES_ONLY, SWB_F_TIE_GT, check


class _Lib:
    def __init__(self, lib):
        self.lib = lib


class _ShardRefs:
    """What AlignResult needs from a RefSet: the library and the shard's sequences."""

    def __init__(self, lib, seqs):
        self.eng = _Lib(lib)
        self.seqs = seqs


class MultiEngine:
    """One process, several devices (swb_multi_*): reference set sharded by length-balanced snake deal,
    every device aligns all reads against its shard on its own host thread, best hits merged with one
    ncclAllGather + a merge kernel."""

    def __init__(self, devices, workspace_bytes: int = 0):
        self.lib = _ffi.load()
        arr = (_C.c_int32 * len(devices))(*devices)
        h = _C.c_void_p()
        check(self.lib.swb_multi_create(arr, len(devices), workspace_bytes, _C.byref(h)))
        self.h, self.n = h, len(devices)
        self.seqs = None

    def close(self):
        if getattr(self, "h", None):
            self.lib.swb_multi_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def load_refset(self, refs):
        data, off, bs = concat(refs)
        check(self.lib.swb_multi_refset_load(self.h, len(bs), data, _i64p(off)))
        self.seqs = bs
        return self

    def shard_refs(self, d: int) -> np.ndarray:
        n = _C.c_int64()
        p = self.lib.swb_multi_shard_refs(self.h, d, _C.byref(n))
        return np.ctypeslib.as_array(p, shape=(n.value,)).copy() if n.value else np.zeros(0, np.int64)

    def ref_location(self, g: int):
        sh, loc = _C.c_int32(), _C.c_int64()
        check(self.lib.swb_multi_ref_location(self.h, g, _C.byref(sh), _C.byref(loc)))
        return sh.value, loc.value

    def align(self, reads, scores=DEFAULT_SCORES, scores_only: bool = False, fetch: bool = True, tie_gt: bool = False):
        data, off, bs = concat(reads)
        flags = (SWB_F_SCORES_ONLY if scores_only else 0) | (0 if fetch else SWB_F_NO_FETCH) | (SWB_F_TIE_GT if tie_gt else 0)
        h = _C.c_void_p()
        check(self.lib.swb_multi_align(self.h, len(bs), data, _i64p(off), scores[0], scores[1], scores[2], flags, _C.byref(h)))
        return MultiResult(self, bs, h, scores_only)


class MultiResult:
    def __init__(self, eng: MultiEngine, reads, h, scores_only):
        self.eng, self.reads, self.h, self.scores_only = eng, reads, h, scores_only
        self.lib = eng.lib
        self._shards = {}

    def free(self):
        for r in self._shards.values():
            r.h = None
        self._shards = {}
        if getattr(self, "h", None):
            self.lib.swb_multi_result_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    @property
    def best_hits(self) -> np.ndarray:
        n = len(self.reads)
        p = self.lib.swb_multi_result_best_hits(self.h)
        return np.ctypeslib.as_array(p, shape=(n * 4,)).copy().reshape(n, 4) if n else np.zeros((0, 4), np.int32)

    @property
    def stats(self) -> dict:
        out = (_C.c_double * 3)()
        check(self.lib.swb_multi_result_stats(self.h, out, 3))
        return {"wall_ms": out[0], "align_ms": out[1], "allgather_merge_ms": out[2]}

    def shard(self, d: int) -> AlignResult:
        """Shard d's own AlignResult (borrowed): local ref index = position in MultiEngine.shard_refs(d)."""
        if d not in self._shards:
            ids = self.eng.shard_refs(d)
            seqs = [self.eng.seqs[int(g)] for g in ids]
            hp = _C.c_void_p(self.lib.swb_multi_result_shard(self.h, d))
            self._shards[d] = AlignResult(_ShardRefs(self.lib, seqs), self.reads, hp, self.scores_only, owned=False).cache()
        return self._shards[d]

    def pair(self, global_ref: int, read: int, max_cells=None):
        sh, loc = self.eng.ref_location(global_ref)
        return self.shard(sh).pair(loc, read, max_cells)


class Comm:
    """One rank of a multi-process job (swb_comm_*): NCCL communicator over this rank's Engine."""

    @staticmethod
    def unique_id() -> bytes:
        buf = _C.create_string_buffer(128)
        check(_ffi.load().swb_comm_unique_id(buf))
        return buf.raw

    def __init__(self, eng, id128: bytes, rank: int, world: int):
        self.lib = _ffi.load()
        h = _C.c_void_p()
        check(self.lib.swb_comm_create(eng.h, id128, rank, world, _C.byref(h)))
        self.h, self.eng = h, eng

    def close(self):
        if getattr(self, "h", None):
            self.lib.swb_comm_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def allgather_best(self, res: AlignResult, global_ids, want_host: bool = True):
        ids = np.ascontiguousarray(global_ids, dtype=np.int64)
        out = np.empty((res.n_reads, 4), dtype=np.int32) if want_host else None
        check(self.lib.swb_comm_allgather_best(self.h, res.h, _i64p(ids), len(ids),
                                               out.ctypes.data_as(_C.POINTER(_C.c_int32)) if want_host else None))
        return out
