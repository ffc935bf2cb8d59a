#!/usr/bin/env python
"""BASELINE config 2 at scale, with sampled parity: N reads x 150 bp against the
10,000 RefSeq-shaped references in ONE swb_align call, then
  * scores of a (reads x refs) sub-block of >= 10^5 pairs against the CPU oracle harness,
  * the per-reference wrapping totals of that block,
  * complete results (cells, beginnings, both strings) of --full-pairs random pairs.

    python tests/checks/run_cfg2_sample.py [--reads 2048] [--score-reads 128] [--score-refs 1000] [--full-pairs 1500]
"""
import argparse
import json
import os
import random
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=2048)
    ap.add_argument("--score-reads", type=int, default=128)
    ap.add_argument("--score-refs", type=int, default=1000)
    ap.add_argument("--full-pairs", type=int, default=1500)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    import numpy as np
    import oracle
    import sparksmithwaterman_b200 as swb
    from sparksmithwaterman_b200 import synth
    refs = synth.make_refs(10_000)
    reads = synth.make_reads(args.reads, 150, refs)
    eng = swb.Engine(0, 16 << 30)
    rs = eng.load_refset(refs)
    t0 = time.perf_counter()
    res = rs.align(reads); res.free()             # first call of this size: grows the device pool and the pinned buffers
    wall_first = time.perf_counter() - t0
    t0 = time.perf_counter()
    res = rs.align(reads)
    wall = time.perf_counter() - t0
    st = res.stats
    scores = res.scores
    out = {"reads": len(reads), "refs": len(refs), "pairs": int(scores.size), "max_cells": res.total_cells,
           "wall_first_call_s": round(wall_first, 3), "wall_s": round(wall, 3), "gcups_e2e": round(st["cells"] / 1e9 / wall, 1),
           "fill_ms": round(st["fill_ms"], 1), "locate_ms": round(st["locate_ms"], 1), "trace_ms": round(st["trace_ms"], 1),
           "d2h_ms": round(st["d2h_ms"], 1), "batches": int(st["batches"])}
    # ---- scores of a sub-block vs the oracle harness --------------------------------------------
    rnd = random.Random(5)
    rsel = sorted(rnd.sample(range(len(refs)), args.score_refs))
    qsel = sorted(rnd.sample(range(len(reads)), args.score_reads))
    cpu = oracle.cpu_baseline([refs[k] for k in rsel], [reads[k] for k in qsel], threads=os.cpu_count() or 1, mode=1,
                              want_scores=True)
    exp = np.array(cpu["pair_scores"], dtype=np.int32).reshape(len(rsel), len(qsel))
    got = scores[np.ix_(rsel, qsel)]
    out["score_pairs_checked"] = int(exp.size)
    out["scores_equal"] = bool((exp == got).all())
    tot = (got.astype(np.int64).sum(axis=1) & 0xFFFFFFFF).astype(np.uint32).view(np.int32)
    out["block_totals_equal"] = bool((tot == np.array(cpu["ref_totals"], dtype=np.int32)).all())
    out["ref_totals_equal_numpy"] = bool((res.ref_totals == (scores.astype(np.int64).sum(axis=1) & 0xFFFFFFFF).astype(np.uint32).view(np.int32)).all())
    out["cpu_gcups"] = round(cpu["gcups"], 3)
    # ---- complete results of random pairs ----------------------------------------------------------
    res.cache()
    ok = True
    for _ in range(args.full_pairs):
        r, q = rnd.randrange(len(refs)), rnd.randrange(len(reads))
        e = oracle.align(refs[r], reads[q])
        g = res.pair(r, q)
        if not (g[0] == e.score and g[1] == e.cells and g[2] == e.sites):
            ok = False
            out["first_mismatch"] = [r, q]
            break
    out["full_pairs_checked"] = args.full_pairs
    out["full_pairs_equal"] = ok
    print(json.dumps(out))
    if args.out:
        with open(args.out, "w") as f:
            json.dump(out, f, indent=1)
    assert out["scores_equal"] and out["block_totals_equal"] and ok and out["ref_totals_equal_numpy"]


if __name__ == "__main__":
    main()
