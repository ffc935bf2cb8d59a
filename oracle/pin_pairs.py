#!/usr/bin/env python
"""Companion of oracle/pin_with_jdk.sh: emits the pinning cases (every known-answer vector of
tests/golden/kat.json + N seeded random pairs over several score sets) as TSV for the Java driver, and compares the
Java operator's answers with the C oracle's (oracle/sw_oracle.c).  The oracle side runs anywhere (gcc only)."""
import argparse
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def cases(n_random):
    with open(os.path.join(ROOT, "tests", "golden", "kat.json")) as f:
        kat = json.load(f)
    out = []
    for v in kat["vectors"] if isinstance(kat, dict) and "vectors" in kat else kat:
        sc = v.get("scores", [5, -3, -4])
        out.append((v["ref"], v["read"], sc[0], sc[1], sc[2]))
    rnd = random.Random(20151004)
    score_sets = [(5, -3, -4), (1, -1, -1), (2, -1, -2), (3, 1, -2), (5, -3, 0), (1, 0, 0), (2, 3, -1)]
    for k in range(n_random):
        alpha = "ACGT" if k % 5 else "AT"
        n, m = rnd.randint(1, 120), rnd.randint(1, 60)
        ref = "".join(rnd.choice(alpha) for _ in range(n))
        read = "".join(rnd.choice(alpha) for _ in range(m))
        if k % 3 == 0:
            ref = ref.lower() if k % 2 else ref
        out.append((ref, read) + score_sets[k % len(score_sets)])
    return [c for c in out if "\t" not in c[0] and "\n" not in c[0] and c[0] and c[1]]


def oracle_line(c):
    import oracle
    r = oracle.align(c[0], c[1], c[2], c[3], c[4], max_cells=64)
    n = len(c[0]) * len(c[1]) if r.score == 0 else None
    if n is None:
        n = len(oracle.align(c[0], c[1], c[2], c[3], c[4]).cells)
    return f"{r.score}\t{n}\t" + "".join(f"{b}:{ra}:{qa};" for (b, ra, qa) in r.sites)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--random", type=int, default=200)
    ap.add_argument("--emit-cases", action="store_true")
    ap.add_argument("--compare", default=None)
    a = ap.parse_args()
    cs = cases(a.random)
    if a.emit_cases:
        for c in cs:
            print("\t".join(str(x) for x in c))
        return
    with open(a.compare, encoding="latin-1") as f:
        java = [l.rstrip("\n") for l in f]
    assert len(java) == len(cs), (len(java), len(cs))
    bad = 0
    for c, j in zip(cs, java):
        o = oracle_line(c)
        if o != j:
            bad += 1
            print("MISMATCH", c, "\n  java  :", j[:200], "\n  oracle:", o[:200])
    print(f"{len(cs) - bad}/{len(cs)} cases equal -- " + ("oracle PINNED to the Java reference" if not bad else "NOT pinned"))
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
