"""Regenerates tests/golden/kat.json.

The reference (Java 8 + Spark) cannot run in this image and ships no golden vectors, so the
vectors are produced by the independent pure-Python twin (oracle/sw_twin.py), which restates
SmithWaterman.java:62-436 without a type matrix; the inputs are the hand-derivable cases of
SURVEY.md section 8c, the EngineerData constants (EngineerData.java:23,26,29) and the
tie-heavy repeats of BASELINE config 5.  tests/test_oracle_golden.py holds the same small
answers hard-coded as well, so a regenerated file cannot silently change them.

    python tests/golden/make_kat.py
"""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import sw_twin  # noqa: E402

REF = "CCTGGGTCCTGCCTCGCATCTGACCAGGGCAGGTGGCCTCCTCATCACACTGCTGCCTCTGCTGTTGGCCCTGCTCATGA"
READ_80 = "AATTTTAGTCTCTCCCTACCCTTTTGGACAGAGCTTCCTGTCCTCTCATTTCACAGGTTATGCAACAGAGGGTTCTGTGT"
READ_20 = "ACTGACTGACTGACTGACTG"

CASES = [
    ("ACGT", "ACGT"), ("AAAA", "CCCC"), ("ACGTACGT", ""), ("", "ACGT"), ("AC", "ACGTT"), ("ATATATAT", "ATAT"),
    ("acgtTTacgt", "ACGT"), ("GATTACA", "GCATGCU"), ("CGTGAATTCAT", "GACTTAC"), ("AAGGAA", "AAA"),
    ("GTTCA", "CTA"), ("CCAAT", "CAT"), ("ATCAA", "TAC"), ("TGGTC", "TGT"),
    (REF * 5, READ_80), (REF * 5, READ_20), (REF * 5, READ_20 * 2), (REF * 5, READ_20 * 5),
    ("AT" * 400, "AT" * 75), ("A" * 300, "A" * 50), ("ACG" * 60, "ACG" * 20), ("AT" * 40 + "G" + "AT" * 40, "AT" * 30),
]
SCORE_SETS = [(5, -3, -4), (1, -1, -1), (2, -2, -2), (3, -3, -1)]


def digest(score, cells, sites):
    h = hashlib.sha256()
    h.update(repr((score, cells, sites)).encode())
    return h.hexdigest()


def main():
    out = []
    for ci, (ref, read) in enumerate(CASES):
        for scores in (SCORE_SETS if ci < 14 else SCORE_SETS[:1]):
            score, cells, sites = sw_twin.align(ref, read, *scores)
            rec = {"ref": ref, "read": read, "scores": list(scores), "score": score, "n_cells": len(cells),
                   "digest": digest(score, [list(c) for c in cells], [list(s) for s in sites])}
            if len(cells) <= 16:
                rec["cells"] = [list(c) for c in cells]
                rec["sites"] = [list(s) for s in sites]
            else:
                rec["cells_head"] = [list(c) for c in cells[:4]]
                rec["cells_tail"] = [list(c) for c in cells[-2:]]
                rec["sites_head"] = [list(s) for s in sites[:2]]
            out.append(rec)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "kat.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=0)
    print(f"wrote {len(out)} vectors to {path}")


if __name__ == "__main__":
    main()
