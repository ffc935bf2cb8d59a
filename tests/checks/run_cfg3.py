#!/usr/bin/env python
"""BASELINE config 3: 10 x (100 kbp vs 100 kbp) pairs through the int32 wide path
(intra-pair banded wavefront + tile-recompute traceback).

Checks, for every pair, size-independent properties of the returned alignment (re-scoring
the 2-bit column string over the original sequences reproduces the score with every prefix
positive; it ends in the reported max cell; `beginning` matches) and bit-exact equality of
EVERY pair (score, max cells, beginnings, both alignment strings by SHA-256) with the CPU
oracle's results committed in tests/golden/cfg3_100k.json (made on the build host by
tests/golden/make_cfg3_golden.py: a minute of CPU per pair that the GPU box is not charged for).
--oracle-pairs N additionally runs the oracle live for the first N pairs.

    python tests/checks/run_cfg3.py [--length 100000] [--pairs 10] [--oracle-pairs 0] [--out profiles/cfg3_r02.json]
"""
import argparse
import json
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


from tests.golden.make_cfg3_golden import cfg3_sequences, site_digest  # noqa: E402


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
    ap.add_argument("--oracle-pairs", type=int, default=0)
    ap.add_argument("--golden", default=None, help="default: tests/golden/cfg3_<length/1000>k.json when it exists")
    ap.add_argument("--repeat", type=int, default=3, help="timed calls (the best is reported)")
    ap.add_argument("--workspace-gb", type=float, default=24.0)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    import sparksmithwaterman_b200 as swb
    L, scores = args.length, (5, -3, -4)
    # 10 pairs = ONE 100 kbp read against 10 references of 100 kbp: 9 homologous (~90 % identity
    # with indels), 1 unrelated -- the reference's pair set is always reads x refs
    read, refs = cfg3_sequences(L, args.pairs)
    gpath = args.golden or os.path.join(ROOT, "tests", "golden", f"cfg3_{L // 1000}k.json")
    golden = None
    if os.path.exists(gpath):
        with open(gpath) as f:
            golden = json.load(f)
        if golden["length"] != L or golden["pairs"] != args.pairs:
            golden = None
    eng = swb.Engine(0, int(args.workspace_gb * (1 << 30)))
    out = {"length": L, "pairs": args.pairs, "scores": scores, "per_pair": [],
           "golden": os.path.relpath(gpath, ROOT) if golden else None}
    rs = eng.load_refset(refs)
    res = rs.align([read])                       # warm-up (pool growth, module load)
    res.free()
    best = None
    for _ in range(max(1, args.repeat)):
        t0 = time.perf_counter()
        res = rs.align([read])
        wall = time.perf_counter() - t0
        st = res.stats
        tt = st["fill_ms"] + st["locate_ms"] + st["trace_ms"]
        if best is None or tt < best[0]:
            if best is not None:
                best[1].free()
            best = (tt, res, wall, st)
        else:
            res.free()
    _, res, wall, st = best
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
        if golden:
            g = golden["per_pair"][k]
            got = res.cache().pair(k, 0)
            rec["golden_equal"] = bool(got[0] == g["score"] and [list(c) for c in got[1]] == g["cells"]
                                       and site_digest(got[2]) == g["sites"])
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
    out["all_golden_equal"] = all(r.get("golden_equal", False) for r in out["per_pair"]) if golden else None
    # int32 roofline of SURVEY 8d (4 integer-pipe ops per cell) and the 2-DPX-op ALU bound of this kernel
    out["roofline_int32_gcups"] = round(148 * 1.965 * 64 / 4, 1)
    out["frac_int32_roofline_fill"] = round(out["gcups_fill"] / (148 * 1.965 * 64 / 4), 3)
    out["frac_alu_bound_fill"] = round(out["gcups_fill"] / (148 * 1.965 * 64 / 2), 3)
    print(json.dumps({k: v for k, v in out.items() if k != "per_pair"}))
    if args.out:
        with open(args.out, "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
