#!/usr/bin/env python
"""BASELINE config 5 at N GPUs: tie-heavy low-complexity repeats.  Every rank aligns all reads
against its reference shard; per-pair results must be bit-exact vs the oracle on every rank's
shard, and the merged best-hit records must be identical for every N (compared with the
single-shard answer computed on rank 0).

    torchrun --nproc-per-node N tests/checks/run_cfg5_multigpu.py
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

import numpy as np
import torch
import torch.distributed as dist


def workload():
    rng = np.random.default_rng(20151005)
    refs = []
    for k in range(48):
        kind = k % 4
        n = int(np.clip(np.rint(np.exp(rng.normal(np.log(600), 0.6))), 60, 4000))
        if kind == 0:
            refs.append("AT" * (n // 2))
        elif kind == 1:
            refs.append("A" * n)
        elif kind == 2:
            refs.append("ACG" * (n // 3))
        else:
            s = list("AT" * (n // 2)); s[len(s) // 3] = "G"; refs.append("".join(s))
    reads = ["AT" * 75, "TA" * 75, "A" * 150, "ACG" * 50, "AT" * 37 + "C" + "AT" * 37, "A" * 70 + "T" + "A" * 79,
             "CCCCCCCCCC" * 15, "GT" * 75]
    return refs, reads


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import oracle
    import sparksmithwaterman_b200 as swb
    from sparksmithwaterman_b200 import multigpu
    refs, reads = workload()
    eng = swb.Engine(local)
    ids = multigpu.shard_refs([len(r) for r in refs], rank, world)
    rs = eng.load_refset([refs[k] for k in ids])
    res = rs.align(reads).cache()
    checked = cells = 0
    for lk, gk in enumerate(ids):
        for q, read in enumerate(reads):
            exp = oracle.align(refs[gk], read)
            got = res.pair(lk, q)
            assert got[0] == exp.score and got[1] == exp.cells and got[2] == exp.sites, (gk, q)
            checked += 1; cells += len(exp.cells)
    # the merge through the C ABI's own collective (swb_comm_*: ncclAllGather + merge kernel), and, as a cross-check,
    # through torch.distributed
    best = torch.from_numpy(multigpu.localize(res.best_hits, ids)).cuda()
    if world > 1:
        box = [multigpu.Comm.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        comm = multigpu.Comm(eng, box[0], rank, world)
        merged = torch.from_numpy(comm.allgather_best(res, np.asarray(ids, dtype=np.int64)))
        assert (merged.numpy() == multigpu.allgather_best_hits(best).cpu().numpy()).all()
        comm.close()
    else:
        merged = best
    # single-shard reference answer from the oracle
    expect = np.zeros((len(reads), 4), np.int32)
    for q, read in enumerate(reads):
        sc = [oracle.score(r, read)[0] for r in refs]
        k = int(np.argmax(sc))
        c = oracle.align(refs[k], read, max_cells=1).cells
        expect[q] = (sc[k], k, c[0][0] if sc[k] > 0 else 0, c[0][1] if sc[k] > 0 else 0)
    ok = bool((merged.cpu().numpy() == expect).all())
    out = {"world": world, "rank": rank, "pairs_checked": checked, "max_cells_checked": cells, "merged_best_hits_equal_single": ok}
    print(json.dumps(out), flush=True)
    assert ok
    if world > 1:
        dist.barrier(); dist.destroy_process_group()


if __name__ == "__main__":
    main()
