#!/usr/bin/env python
"""BASELINE config 2 at its stated size: 100,000 x 150 bp reads against 10,000 RefSeq-shaped references on ONE
B200 (3.2e14 cells, 1e9 pairs), fed in chunks through the reference-facing C-ABI call with HOST buffers
(swb_align: H2D of the reads, fill, every max cell, every traceback, D2H of every result array).

Parity while it runs (the oracle works on host threads beside the GPU): per chunk a random sample of pair scores and
a few complete pairs (cells, beginnings, both strings) against the CPU oracle, the per-reference wrapping totals
against the chunk's own score matrix, and at the end a checksum of checksums over all 1e9 scores.

Chunk k's reads are synth.make_reads(chunk, 150, refs, seed = READ_SEED + k): deterministic, generated on a host
thread while the previous chunk is on the GPU.

    python tests/checks/run_cfg2_full.py [--reads 100000] [--chunk 2048] [--out profiles/cfg2_full_r02.json]
"""
import argparse
import json
import os
import random
import sys
import time
from concurrent.futures import ThreadPoolExecutor

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=100_000)
    ap.add_argument("--chunk", type=int, default=2048)
    ap.add_argument("--refs", type=int, default=10_000)
    ap.add_argument("--workspace-gb", type=float, default=64.0)
    ap.add_argument("--scores-per-chunk", type=int, default=2048)
    ap.add_argument("--full-per-chunk", type=int, default=24)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    import numpy as np
    import oracle
    import sparksmithwaterman_b200 as swb
    from sparksmithwaterman_b200 import synth
    oracle.build()
    refs = synth.make_refs(a.refs)
    ref_bases = sum(len(r) for r in refs)
    eng = swb.Engine(0, int(a.workspace_gb * (1 << 30)))
    t0 = time.perf_counter()
    rs = eng.load_refset(refs)
    load_s = time.perf_counter() - t0
    sizes = [min(a.chunk, a.reads - k) for k in range(0, a.reads, a.chunk)]
    gen = ThreadPoolExecutor(1)
    chk = ThreadPoolExecutor(max(2, (os.cpu_count() or 4) - 2))
    make = lambda k: synth.make_reads(sizes[k], 150, refs, seed=synth.READ_SEED + k)
    nxt = gen.submit(make, 0)
    # warm-up on a throw-away chunk: pool growth, pinned result buffers
    res = rs.align(synth.make_reads(a.chunk, 150, refs, seed=synth.READ_SEED - 1)); res.free()
    rnd = random.Random(2)
    pending, t_align, n_cells_total, n_pairs, checksum, fill_ms, trace_ms, locate_ms, d2h_ms = [], 0.0, 0, 0, 0, 0.0, 0.0, 0.0, 0.0
    ok_totals = True
    t_wall0 = time.perf_counter()
    for k in range(len(sizes)):
        reads = nxt.result()
        if k + 1 < len(sizes):
            nxt = gen.submit(make, k + 1)
        t0 = time.perf_counter()
        res = rs.align(reads)                                  # host buffers in, every result array out
        t_align += time.perf_counter() - t0
        st = res.stats
        fill_ms += st["fill_ms"]; trace_ms += st["trace_ms"]; locate_ms += st["locate_ms"]; d2h_ms += st["d2h_ms"]
        n_cells_total += res.total_cells
        sc = res.scores
        n_pairs += sc.size
        tot = (sc.astype(np.int64).sum(axis=1) & 0xFFFFFFFF).astype(np.uint32).view(np.int32)
        ok_totals &= bool((res.ref_totals == tot).all())
        checksum = (checksum * 1000003 + int(sc.astype(np.int64).sum()) + k) & 0xFFFFFFFFFFFFFFFF     # checksum of checksums
        res.cache()
        for _ in range(a.scores_per_chunk):
            r, q = rnd.randrange(len(refs)), rnd.randrange(len(reads))
            pending.append(chk.submit(lambda r=r, rd=reads[q], got=int(sc[r, q]): oracle.score(refs[r], rd)[0] == got))
        for _ in range(a.full_per_chunk):
            r, q = rnd.randrange(len(refs)), rnd.randrange(len(reads))
            got = res.pair(r, q)
            def full(r=r, rd=reads[q], got=got):
                e = oracle.align(refs[r], rd)
                return got[0] == e.score and got[1] == e.cells and got[2] == e.sites
            pending.append(chk.submit(full))
        res.free()
        if k % 8 == 0:
            print(f"chunk {k + 1}/{len(sizes)}: {t_align:.1f} s in swb_align so far", flush=True)
    t_wall = time.perf_counter() - t_wall0
    results = [p.result() for p in pending]
    cells = ref_bases * 150 * a.reads
    out = {"config": "cfg2 at stated size", "reads": a.reads, "refs": a.refs, "ref_bases": ref_bases, "pairs": n_pairs, "cells": cells,
           "chunk_reads": a.chunk, "chunks": len(sizes), "refset_load_s": round(load_s, 3),
           "seconds_in_swb_align": round(t_align, 2), "wall_s_incl_host_generation_and_checks": round(t_wall, 2),
           "gcups_e2e_host_buffers": round(cells / 1e9 / t_align, 1), "reads_per_s": round(a.reads / t_align, 1),
           "fill_ms": round(fill_ms, 1), "locate_ms": round(locate_ms, 1), "trace_ms": round(trace_ms, 1), "d2h_ms": round(d2h_ms, 1),
           "gcups_fill_only": round(cells / 1e9 / (fill_ms * 1e-3), 1),
           "max_cells_total": n_cells_total, "score_checksum_of_checksums": checksum,
           "sampled_scores_checked": a.scores_per_chunk * len(sizes), "full_pairs_checked": a.full_per_chunk * len(sizes),
           "all_sampled_equal_to_oracle": bool(all(results)), "ref_totals_equal_score_sums": ok_totals}
    print(json.dumps(out))
    if a.out:
        with open(a.out, "w") as f:
            json.dump(out, f, indent=1)
    assert out["all_sampled_equal_to_oracle"] and ok_totals
    rs.free(); eng.close()


if __name__ == "__main__":
    main()
