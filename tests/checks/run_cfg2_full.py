#!/usr/bin/env python
"""BASELINE config 2 at its stated size: 100,000 x 150 bp reads against 10,000 RefSeq-shaped references on ONE
B200 (3.2e14 cells, 1e9 pairs), fed in chunks through the reference-facing C-ABI call with HOST buffers
(swb_align: H2D of the reads, fill, every max cell, every traceback, D2H of every result array).

Parity: per chunk a random sample of pair scores and
a few complete pairs (cells, beginnings, both strings) against the CPU oracle, the per-reference wrapping totals
against the chunk's own score matrix, and at the end a checksum of checksums over all 1e9 scores.

Chunk k's reads are synth.make_reads(chunk, 150, refs, seed = READ_SEED + k): deterministic, generated in a worker
process while the previous chunk is on the GPU.

    python tests/checks/run_cfg2_full.py [--reads 100000] [--chunk 2048] [--out profiles/cfg2_full_r02.json]
"""
import argparse
import json
import os
import random
import sys
import time
from concurrent.futures import ProcessPoolExecutor, ThreadPoolExecutor

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))


_REFS = None


def _make_chunk(n, seed):
    from sparksmithwaterman_b200 import synth
    return synth.make_reads(n, 150, _REFS, seed=seed)


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
    import multiprocessing
    import numpy as np
    import oracle
    from sparksmithwaterman_b200 import synth
    oracle.build()
    refs = synth.make_refs(a.refs)
    ref_bases = sum(len(r) for r in refs)
    sizes = [min(a.chunk, a.reads - k) for k in range(0, a.reads, a.chunk)]
    # The next chunk's reads are generated in a worker PROCESS (forked before CUDA is initialised; it inherits the
    # references) and the oracle checks run after the timed loop: Python threads working beside swb_align would make
    # the calling thread wait for the interpreter lock and that wait would be charged to the engine.
    global _REFS
    _REFS = refs
    gen = ProcessPoolExecutor(1, mp_context=multiprocessing.get_context("fork"))
    nxt = gen.submit(_make_chunk, sizes[0], synth.READ_SEED)
    import sparksmithwaterman_b200 as swb
    eng = swb.Engine(0, int(a.workspace_gb * (1 << 30)))
    t0 = time.perf_counter()
    rs = eng.load_refset(refs)
    load_s = time.perf_counter() - t0
    # warm-up on a throw-away chunk: pool growth, pinned result buffers
    res = rs.align(synth.make_reads(a.chunk, 150, refs, seed=synth.READ_SEED - 1)); res.free()
    rnd = random.Random(2)
    todo, t_align, n_cells_total, n_pairs, checksum, fill_ms, trace_ms, locate_ms, d2h_ms = [], 0.0, 0, 0, 0, 0.0, 0.0, 0.0, 0.0
    ok_totals = True
    t_wall0 = time.perf_counter()
    for k in range(len(sizes)):
        reads = nxt.result()
        if k + 1 < len(sizes):
            nxt = gen.submit(_make_chunk, sizes[k + 1], synth.READ_SEED + k + 1)
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
            todo.append((r, reads[q], int(sc[r, q]), None))
        for _ in range(a.full_per_chunk):
            r, q = rnd.randrange(len(refs)), rnd.randrange(len(reads))
            todo.append((r, reads[q], None, res.pair(r, q)))
        res.free()
        if k % 8 == 0:
            print(f"chunk {k + 1}/{len(sizes)}: {t_align:.1f} s in swb_align so far", flush=True)
    t_wall = time.perf_counter() - t_wall0
    gen.shutdown()

    def check(item):
        r, rd, score, full = item
        if full is None:
            return oracle.score(refs[r], rd)[0] == score
        e = oracle.align(refs[r], rd)
        return full[0] == e.score and full[1] == e.cells and full[2] == e.sites
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max(2, (os.cpu_count() or 4))) as chk:      # the oracle's C calls release the interpreter lock
        results = list(chk.map(check, todo))
    check_s = time.perf_counter() - t0
    cells = ref_bases * 150 * a.reads
    out = {"config": "cfg2 at stated size", "reads": a.reads, "refs": a.refs, "ref_bases": ref_bases, "pairs": n_pairs, "cells": cells,
           "chunk_reads": a.chunk, "chunks": len(sizes), "refset_load_s": round(load_s, 3),
           "seconds_in_swb_align": round(t_align, 2), "wall_s_of_the_loop_incl_result_sampling": round(t_wall, 2), "oracle_check_s_afterwards": round(check_s, 2),
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
