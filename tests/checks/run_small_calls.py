#!/usr/bin/env python
"""The UNCHANGED-driver path: the reference's MapRef calls `new SmithWaterman.OptAlignments().call({ref, read}, ...)`
once per pair from N Spark task threads (Distribution.java:419-426).  Served one by one each such call is
one native round trip: swb_refset_load(1 ref) + swb_align(1 read) + accessors + frees ("per_pair"); through the
submission queue (swb_align_pair, "per_pair_queued") concurrent calls are coalesced.  This measures pairs/s of both
with N host threads sharing ONE context (ctypes releases the GIL during the native calls), the per-reference
batched call of the changed MapRef (1 ref x R reads), and the per-file batched call, on the same pairs.

    python tests/checks/run_small_calls.py [--threads 1,4,16] [--out profiles/small_calls_r02.json]
"""
import argparse
import json
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--threads", default="1,4,16,64")
    ap.add_argument("--refs", type=int, default=256)
    ap.add_argument("--reads", type=int, default=16)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    import sparksmithwaterman_b200 as swb
    from sparksmithwaterman_b200 import synth
    refs = synth.make_refs(a.refs)
    reads = synth.make_reads(a.reads, 150, refs)
    eng = swb.Engine(0)
    pairs = [(r, q) for r in range(len(refs)) for q in range(len(reads))]

    def one_pair(rq):
        r, q = rq
        rs = eng.load_refset([refs[r]])
        res = rs.align([reads[q]])
        out = res.pair(0, 0)
        res.free(); rs.free()
        return out[0]

    def one_pair_queued(rq):
        r, q = rq
        res = eng.align_pair(refs[r], reads[q])
        out = res.pair(0, 0)
        res.free()
        return out[0]

    def one_ref(r):
        rs = eng.load_refset([refs[r]])
        res = rs.align(reads).cache()
        out = [res.pair(0, q)[0] for q in range(len(reads))]
        res.free(); rs.free()
        return out

    out = {"refs": len(refs), "reads": len(reads), "pairs": len(pairs), "per_pair": [], "per_pair_queued": [], "per_ref": []}
    for n in [int(x) for x in a.threads.split(",")]:
        with ThreadPoolExecutor(n) as ex:
            list(ex.map(one_pair, pairs[:64]))                         # warm
            t0 = time.perf_counter()
            sc = list(ex.map(one_pair, pairs[:2048]))
            dt = time.perf_counter() - t0
            out["per_pair"].append({"threads": n, "pairs_per_s": round(len(sc) / dt, 1), "us_per_pair": round(dt / len(sc) * 1e6, 1)})
            list(ex.map(one_pair_queued, pairs[:64]))
            q0 = eng.queue_stats()
            t0 = time.perf_counter()
            sq = list(ex.map(one_pair_queued, pairs[:2048]))
            dt = time.perf_counter() - t0
            q1 = eng.queue_stats()
            assert sq == sc
            out["per_pair_queued"].append({"threads": n, "pairs_per_s": round(len(sq) / dt, 1), "us_per_pair": round(dt / len(sq) * 1e6, 1),
                                           "batches": q1["batches"] - q0["batches"], "largest_batch": q1["largest_batch"]})
            t0 = time.perf_counter()
            rows = list(ex.map(one_ref, range(len(refs))))
            dt = time.perf_counter() - t0
            out["per_ref"].append({"threads": n, "pairs_per_s": round(len(refs) * len(reads) / dt, 1), "us_per_call": round(dt / len(refs) * 1e6, 1)})
    rs = eng.load_refset(refs)
    rs.align(reads).free()
    t0 = time.perf_counter()
    res = rs.align(reads).cache()
    got = [[res.pair(r, q)[0] for q in range(len(reads))] for r in range(len(refs))]
    dt = time.perf_counter() - t0
    out["per_file"] = {"pairs_per_s": round(len(pairs) / dt, 1), "ms_per_call": round(dt * 1e3, 2)}
    assert got == rows
    out["per_pair_vs_per_file"] = round(out["per_pair"][-1]["pairs_per_s"] / out["per_file"]["pairs_per_s"], 5)
    out["per_pair_queued_vs_per_file"] = round(max(x["pairs_per_s"] for x in out["per_pair_queued"]) / out["per_file"]["pairs_per_s"], 5)
    print(json.dumps(out))
    if a.out:
        with open(a.out, "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
