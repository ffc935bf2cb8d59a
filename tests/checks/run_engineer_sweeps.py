#!/usr/bin/env python
"""Runs the reference's four benchmark sweeps (datasets.py) through the engine on cuda:0 and
writes one record per point: wall ms of swb_align (host buffers in, results out), GCUPS, max
score; a sample of points is checked pair-by-pair against the CPU oracle.

    python tests/checks/run_engineer_sweeps.py [--out profiles/engineer_sweeps_r01.json] [--check-every 6]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--check-every", type=int, default=6)
    ap.add_argument("--sweeps", default="read_num,read_len,ref_num,ref_len")
    args = ap.parse_args()
    import oracle
    import sparksmithwaterman_b200 as swb
    from sparksmithwaterman_b200 import datasets
    eng = swb.Engine(0)
    out = []
    for name in args.sweeps.split(","):
        for k, (sweep, x, refs, reads) in enumerate(datasets.SWEEPS[name]()):
            rs = eng.load_refset(refs)
            res = rs.align(reads); res.free()                       # warm
            t0 = time.perf_counter()
            res = rs.align(reads)
            ms = (time.perf_counter() - t0) * 1e3
            st = res.stats
            rec = {"sweep": sweep, "x": x, "refs": len(refs), "reads": len(reads), "ms": round(ms, 3),
                   "device_ms": round(st["device_ms"], 3), "gcups": round(st["cells"] / 1e9 / (ms * 1e-3), 1),
                   "max_score": int(res.scores.max()), "max_cells": int(res.total_cells)}
            if k % args.check_every == 0:
                # distinct (ref, read) pairs only: the sweeps repeat the same sequences
                seen = {}
                res.cache()
                for r, ref in enumerate(refs):
                    for q, read in enumerate(reads):
                        if (ref, read) in seen:
                            continue
                        seen[(ref, read)] = 1
                        exp = oracle.align(ref, read)
                        got = res.pair(r, q)
                        assert got[0] == exp.score and got[1] == exp.cells and got[2] == exp.sites, (sweep, x)
                rec["oracle_checked_pairs"] = len(seen)
            out.append(rec)
            print(json.dumps(rec), flush=True)
            res.free(); rs.free()
    if args.out:
        with open(args.out, "w") as f:
            json.dump(out, f, indent=0)


if __name__ == "__main__":
    main()
