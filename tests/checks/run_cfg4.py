#!/usr/bin/env python
"""BASELINE config 4: 150 bp reads against a 1 % RefSeq-shaped set (123,212 references, ~266 Mbp) sharded over
the GPUs of one box -- one process per GPU (torchrun), reference shards resident in HBM, every rank aligns all
reads against its shard through swb_align (host buffers), the per-read best-hit records are merged by libswb200's
own ncclAllGather + merge kernel (swb_comm_*).  The stated job is 1M reads; the run does --reads of them
(GCUPS is size-invariant once saturated) in chunks.

Afterwards rank 0 loads the WHOLE set on its GPU and recomputes a sample of the reads on a single device: the merged
records must be identical.

    torchrun --nproc-per-node 8 tests/checks/run_cfg4.py [--reads 16384] [--chunk 2048] [--out profiles/cfg4_r02.json]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

import numpy as np
import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--refs", type=int, default=123_212)
    ap.add_argument("--reads", type=int, default=16_384)
    ap.add_argument("--chunk", type=int, default=2048)
    ap.add_argument("--sample", type=int, default=48)
    ap.add_argument("--workspace-gb", type=float, default=64.0)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import sparksmithwaterman_b200 as swb
    from sparksmithwaterman_b200 import multigpu, synth
    refs = synth.make_refs(a.refs)
    ids = multigpu.shard_refs([len(r) for r in refs], rank, world)
    ids_np = np.asarray(ids, dtype=np.int64)
    eng = swb.Engine(local, int(a.workspace_gb * (1 << 30)))
    t0 = time.perf_counter()
    rs = eng.load_refset([refs[k] for k in ids])
    torch.cuda.synchronize()
    load_s = time.perf_counter() - t0
    comm = None
    if world > 1:
        box = [multigpu.Comm.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        comm = multigpu.Comm(eng, box[0], rank, world)
    sizes = [min(a.chunk, a.reads - k) for k in range(0, a.reads, a.chunk)]
    # all chunks up front (same bytes on every rank): host generation is not part of the job
    chunks = [synth.make_reads(n, 150, refs[:4096], seed=synth.READ_SEED + k) for k, n in enumerate(sizes)]
    res = rs.align(synth.make_reads(sizes[0], 150, refs[:4096], seed=synth.READ_SEED - 1)); res.free()   # warm-up on a throw-away chunk of the real size: pinned result buffers, device pool
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    merged_all, cells_max = [], 0
    t0 = time.perf_counter()
    for reads in chunks:
        res = rs.align(reads)
        cells_max += res.total_cells
        if comm:
            merged_all.append(comm.allgather_best(res, ids_np, want_host=True))
        else:
            merged_all.append(multigpu.localize(res.best_hits, ids))
        res.free()
    torch.cuda.synchronize()
    secs = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([secs, load_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        secs, load_max = float(t[0]), float(t[1])
        c = torch.tensor([float(cells_max)], device="cuda", dtype=torch.float64)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        cells_max = int(c.item())
    else:
        load_max = load_s
    merged = np.concatenate(merged_all)
    ok = None
    sample_s = None
    if rank == 0:
        # single-device recomputation of a sample of the reads against the WHOLE set
        rs.free()
        t1 = time.perf_counter()
        rs_all = eng.load_refset(refs)
        load_all_s = time.perf_counter() - t1
        flat = [r for ch in chunks for r in ch]
        pick = sorted(np.random.default_rng(4).choice(len(flat), size=min(a.sample, len(flat)), replace=False).tolist())
        t1 = time.perf_counter()
        one = rs_all.align([flat[k] for k in pick])
        sample_s = time.perf_counter() - t1
        ok = bool((one.best_hits == merged[pick]).all())
        one.free(); rs_all.free()
        ref_bases = sum(len(r) for r in refs)
        cells = ref_bases * 150 * a.reads
        out = {"config": "cfg4: 150 bp reads vs 123,212 RefSeq-shaped refs sharded over the box", "n_gpus": world, "refs": a.refs,
               "ref_bases": ref_bases, "refs_per_gpu": len(ids), "reads_run": a.reads, "reads_stated": 1_000_000, "chunk_reads": a.chunk,
               "cells": cells, "seconds": round(secs, 3), "gcups_whole_box_e2e_host_buffers": round(cells / 1e9 / secs, 1),
               "reads_per_s": round(a.reads / secs, 1), "projected_s_for_1M_reads": round(1e6 / (a.reads / secs), 1),
               "shard_load_s_max": round(load_max, 3), "whole_set_load_s_one_gpu": round(load_all_s, 3),
               "max_cells": cells_max, "allgather_bytes_per_rank_per_chunk": 16 * a.chunk,
               "collective": "libswb200 swb_comm_allgather_best (ncclAllGather + merge kernel on the engine stream)",
               "sample_reads_recomputed_on_one_gpu": len(pick), "sample_seconds": round(sample_s, 2),
               "merged_best_hits_equal_single_gpu": ok}
        print(json.dumps(out), flush=True)
        if a.out:
            with open(a.out, "w") as f:
                json.dump(out, f, indent=1)
    if comm:
        comm.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        assert ok


if __name__ == "__main__":
    main()
