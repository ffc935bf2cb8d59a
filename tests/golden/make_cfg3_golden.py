#!/usr/bin/env python
"""Golden results of BASELINE config 3 (10 x (100 kbp vs 100 kbp) pairs) from the CPU oracle.

The oracle (oracle/sw_oracle.c, two-row scores + 2-bit-plane traceback; a restatement of
SmithWaterman.java:62-436) needs about a minute and 2.5 GB per pair, which is CPU time the GPU box should not
be charged for: this script runs it once on the build host and commits, per pair, the score, the
maximum cells, the beginnings and SHA-256 digests of the two alignment strings.  The GPU checkers
(tests/checks/run_cfg3.py, tests/test_gpu_cfg_sizes.py) rebuild the same seeded sequences with
cfg3_sequences() and compare.

    python tests/golden/make_cfg3_golden.py --length 100000 --out tests/golden/cfg3_100k.json
    python tests/golden/make_cfg3_golden.py --length 20000 --pairs 4 --out tests/golden/cfg3_20k.json
"""
import argparse
import hashlib
import json
import os
import random
import sys
from concurrent.futures import ProcessPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

SEED = 20151003
SCORES = (5, -3, -4)


def mutate(rnd, s, sub, indel):
    out = []
    for ch in s:
        u = rnd.random()
        if u < indel / 2:
            continue
        if u < indel:
            out.append(rnd.choice("ACGT"))
        out.append(rnd.choice("ACGT") if rnd.random() < sub else ch)
    return "".join(out)


def cfg3_sequences(length, pairs, seed=SEED):
    """ONE read of `length` bp against `pairs` references: pairs-1 homologous (~90 % identity with indels),
    1 unrelated -- the reference's pair set is always reads x refs (Distribution.java:714-724)."""
    rnd = random.Random(seed)
    read = "".join(rnd.choice("ACGT") for _ in range(length))
    refs = [mutate(rnd, read, 0.07, 0.03)[:length] for _ in range(pairs - 1)]
    refs.append("".join(rnd.choice("ACGT") for _ in range(length)))
    return read, refs


def site_digest(sites):
    """[(beginning, sha256(ref_aln), sha256(read_aln), len)] of a pair's alignments."""
    return [[int(b), hashlib.sha256(ra.encode()).hexdigest(), hashlib.sha256(qa.encode()).hexdigest(), len(ra)]
            for (b, ra, qa) in sites]


def _one(args):
    k, ref, read = args
    import oracle
    exp = oracle.align(ref, read, *SCORES, lowmem=True)
    return {"pair": k, "score": exp.score, "cells": [list(c) for c in exp.cells], "sites": site_digest(exp.sites)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--length", type=int, default=100_000)
    ap.add_argument("--pairs", type=int, default=10)
    ap.add_argument("--procs", type=int, default=5)
    ap.add_argument("--out", required=True)
    a = ap.parse_args()
    read, refs = cfg3_sequences(a.length, a.pairs)
    import oracle
    oracle.build()
    with ProcessPoolExecutor(a.procs) as ex:
        per = list(ex.map(_one, [(k, refs[k], read) for k in range(a.pairs)]))
    out = {"what": "oracle results of cfg3_sequences(length, pairs, seed); made by tests/golden/make_cfg3_golden.py",
           "length": a.length, "pairs": a.pairs, "seed": SEED, "scores": SCORES, "per_pair": per}
    with open(a.out, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", a.out, [p["score"] for p in per])


if __name__ == "__main__":
    main()
