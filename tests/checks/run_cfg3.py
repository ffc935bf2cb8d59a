#!/usr/bin/env python
"""BASELINE config 3: 10 x (100 kbp vs 100 kbp) pairs through the int32 wide path
(intra-pair banded wavefront + tile-recompute traceback).

Checks, for every pair, size-independent properties of the returned alignment (re-scoring
the 2-bit column string over the original sequences reproduces the score with every prefix
positive; it ends in the reported max cell; `beginning` matches) and, for --oracle-pairs
pairs, bit-exact equality with the CPU oracle (two-row score + 2-bit-plane traceback).

    python tests/checks/run_cfg3.py [--length 100000] [--pairs 10] [--oracle-pairs 1] [--out profiles/cfg3_r01.json]
"""
import argparse
import json
import os
import random
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))


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


def rescore(ref, read, i, j, ops, scores):
    """Walk the columns backwards from (i, j); return (score, start column, ok)."""
    match, mismatch, gap = scores
    total, ok = 0, True
    suffix = []
    for op in reversed(ops):
        if op == 1:
            total += match if ref[j - 1].upper() == read[i - 1].upper() else mismatch
            i -= 1; j -= 1
        elif op == 2:
            total += gap; i -= 1
        else:
            total += gap; j -= 1
        suffix.append(total)
    # every prefix of the forward path must be positive: prefix_k = total - suffix_(len-k)
    for k in range(len(suffix) - 1):
        if total - suffix[k] <= 0:
            ok = False
            break
    return total, j + 1, ok and i >= 0 and j >= 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--length", type=int, default=100_000)
    ap.add_argument("--pairs", type=int, default=10)
    ap.add_argument("--oracle-pairs", type=int, default=1)
    ap.add_argument("--workspace-gb", type=float, default=24.0)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    import sparksmithwaterman_b200 as swb
    rnd = random.Random(20151003)
    L, scores = args.length, (5, -3, -4)
    # 10 pairs = ONE 100 kbp read against 10 references of 100 kbp: 9 homologous (~90 % identity
    # with indels), 1 unrelated -- the reference's pair set is always reads x refs
    read = "".join(rnd.choice("ACGT") for _ in range(L))
    refs = [mutate(rnd, read, 0.07, 0.03)[:L] for _ in range(args.pairs - 1)]
    refs.append("".join(rnd.choice("ACGT") for _ in range(L)))
    eng = swb.Engine(0, int(args.workspace_gb * (1 << 30)))
    out = {"length": L, "pairs": args.pairs, "scores": scores, "per_pair": []}
    rs = eng.load_refset(refs)
    res = rs.align([read])                       # warm-up (pool growth, module load)
    res.free()
    t0 = time.perf_counter()
    res = rs.align([read])
    wall = time.perf_counter() - t0
    st = res.stats
    cells = res.cells; begs = res.beginnings; offs = res.cell_offsets
    for k in range(args.pairs):
        score = int(res.scores[k, 0])
        n_cells = res.pair_cell_count(k, 0)
        rec = {"pair": k, "m": len(read), "n": len(refs[k]), "score": score, "max_cells": n_cells}
        ok_all = True
        for c in range(int(offs[k]), int(offs[k]) + min(n_cells, 4)):
            ops = res.ops(c)
            total, start, ok = rescore(refs[k], read, int(cells[c][0]), int(cells[c][1]), ops, scores)
            ok_all &= ok and total == score and start == int(begs[c])
            rec.setdefault("aln_len", []).append(int(len(ops)))
        rec["rescore_ok"] = bool(ok_all)
        if k < args.oracle_pairs:
            import oracle
            t1 = time.perf_counter()
            exp = oracle.align(refs[k], read, *scores, lowmem=True)
            rec["oracle_s"] = round(time.perf_counter() - t1, 1)
            got = res.cache().pair(k, 0)
            rec["oracle_equal"] = bool(got[0] == exp.score and got[1] == exp.cells and got[2] == exp.sites)
        out["per_pair"].append(rec)
        print(json.dumps(rec), flush=True)
    cells_total = sum(len(r) for r in refs) * len(read)
    t_total = st["fill_ms"] + st["locate_ms"] + st["trace_ms"]
    out.update({"fill_ms": round(st["fill_ms"], 2), "locate_ms": round(st["locate_ms"], 2), "trace_ms": round(st["trace_ms"], 2),
                "d2h_ms": round(st["d2h_ms"], 2), "wall_ms": round(wall * 1e3, 2), "batches": int(st["batches"]),
                "workspace_bytes": int(st["checkpoint_bytes"])})
    out["gcups_fill"] = round(cells_total / 1e9 / (st["fill_ms"] * 1e-3), 1)
    out["gcups_incl_traceback"] = round(cells_total / 1e9 / (t_total * 1e-3), 1)
    out["all_rescore_ok"] = all(r["rescore_ok"] for r in out["per_pair"])
    print(json.dumps({k: v for k, v in out.items() if k != "per_pair"}))
    if args.out:
        with open(args.out, "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
