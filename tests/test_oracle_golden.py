"""CPU: the C oracle against (1) hand-derivable known answers of SURVEY.md section 8c,
hard-coded here, (2) the committed golden vectors (tests/golden/kat.json, produced by the
independent Python twin), (3) the twin itself on random + adversarial inputs."""
import hashlib
import json
import os
import random

import pytest
from hypothesis import given, settings, strategies as st

import oracle
from oracle import sw_twin

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "CCTGGGTCCTGCCTCGCATCTGACCAGGGCAGGTGGCCTCCTCATCACACTGCTGCCTCTGCTGTTGGCCCTGCTCATGA"
READ_80 = "AATTTTAGTCTCTCCCTACCCTTTTGGACAGAGCTTCCTGTCCTCTCATTTCACAGGTTATGCAACAGAGGGTTCTGTGT"
READ_20 = "ACTGACTGACTGACTGACTG"

# (ref, read) -> (score, cells, [(beginning, refAln, readAln)])  -- scores 5/-3/-4
HAND = [
    ("ACGT", "ACGT", 20, [(4, 4)], [(1, "ACGT", "ACGT")]),
    ("ACGTACGT", "", 0, [], []),
    ("AC", "ACGTT", 10, [(2, 2)], [(1, "AC", "AC")]),
    ("ATATATAT", "ATAT", 20, [(4, 4), (4, 6), (4, 8)], [(1, "ATAT", "ATAT"), (3, "ATAT", "ATAT"), (5, "ATAT", "ATAT")]),
    ("acgtTTacgt", "ACGT", 20, [(4, 4), (4, 10)], [(1, "acgt", "ACGT"), (7, "acgt", "ACGT")]),
    ("GATTACA", "GCATGCU", 11, [(4, 3)], [(1, "G_AT", "GCAT")]),
    ("CGTGAATTCAT", "GACTTAC", 18, [(6, 10), (7, 9)], [(4, "GAATTCA", "GACTT_A"), (4, "GAATT_C", "GACTTAC")]),
    ("AAGGAA", "AAA", 10, [(2, 2), (2, 6), (3, 2), (3, 6)], [(1, "AA", "AA"), (5, "AA", "AA"), (1, "AA", "AA"), (5, "AA", "AA")]),
    ("GTTCA", "CTA", 6, [(3, 5)], [(4, "C_A", "CTA")]),          # ">" instead of ">=" would give (3, TCA, T_A)
    ("CCAAT", "CAT", 11, [(3, 5)], [(2, "CAAT", "C_AT")]),       # ">" would give CA_T
    ("ATCAA", "TAC", 6, [(2, 4), (3, 3)], [(2, "TCA", "T_A"), (2, "T_C", "TAC")]),
    ("TGGTC", "TGT", 11, [(3, 4)], [(1, "TGGT", "T_GT")]),
]


@pytest.mark.parametrize("ref,read,score,cells,sites", HAND)
def test_hand_derived(ref, read, score, cells, sites):
    for lowmem in (False, True):
        r = oracle.align(ref, read, lowmem=lowmem)
        assert (r.score, r.cells, r.sites) == (score, cells, sites)
    assert sw_twin.align(ref, read) == (score, cells, sites)


def test_all_zero_matrix_lists_every_cell():
    r = oracle.align("AAAA", "CCCC")
    assert r.score == 0
    assert r.cells == [(i, j) for i in range(1, 5) for j in range(1, 5)]
    assert r.sites == [(0, "", "")] * 16


def test_engineer_data_vectors():
    r = oracle.align(REF * 5, READ_80)
    assert r.score == 124 and r.cells == [(78, 75), (78, 155), (78, 235), (78, 315), (78, 395)]
    assert [s[0] for s in r.sites] == [6, 86, 166, 246, 326]
    assert r.sites[0][1] == "GTC_CTGCCT_CGCATCT_GACCAGGGC_AGGTGGCCTCCTCA__TCACA_CTGCTGCCTCTGCTGTTGGCCCTGCT"
    assert r.sites[0][2] == "GTCTCTCCCTACCCTTTTGGA_CAGAGCTTCCTGTCCT_CTCATTTCACAGGTTATGCAACAG_AG__GGTTCTG_T"
    r = oracle.align(REF * 5, READ_20)
    assert r.score == 56 and [c[1] for c in r.cells] == [64, 144, 224, 304, 384]
    assert r.sites[0] == (49, "ACTG_CTGCCT__CTG_CTG", "ACTGACTGACTGACTGACTG")
    assert oracle.align(REF * 5, READ_20 * 2).score == 93
    r = oracle.align(REF * 5, READ_20 * 5)
    assert r.score == 190 and r.cells == [(100, 144), (100, 224), (100, 304), (100, 384)]


def test_tie_heavy_vectors():
    r = oracle.align("AT" * 400, "AT" * 75)
    assert r.score == 750 and len(r.cells) == 326
    assert r.cells[0] == (150, 150) and r.cells[-1] == (150, 800)
    assert all(len(s[1]) == 150 for s in r.sites) and [s[0] for s in r.sites[:3]] == [1, 3, 5]
    r = oracle.align("A" * 300, "A" * 50)
    assert r.score == 250 and len(r.cells) == 251 and all(len(s[1]) == 50 for s in r.sites)


def _digest(score, cells, sites):
    h = hashlib.sha256()
    h.update(repr((score, [list(c) for c in cells], [list(s) for s in sites])).encode())
    return h.hexdigest()


def test_golden_file():
    with open(os.path.join(HERE, "golden", "kat.json")) as f:
        vectors = json.load(f)
    assert len(vectors) >= 60
    for v in vectors:
        r = oracle.align(v["ref"], v["read"], *v["scores"])
        assert r.score == v["score"] and len(r.cells) == v["n_cells"], (v["ref"][:20], v["read"][:20])
        assert _digest(r.score, r.cells, r.sites) == v["digest"]
        if "cells" in v:
            assert [list(c) for c in r.cells] == v["cells"]
            assert [list(s) for s in r.sites] == v["sites"]


seqs = st.text(alphabet="ACGTacgtN", min_size=0, max_size=40)
low_complexity = st.text(alphabet="AT", min_size=0, max_size=40)
score_sets = st.sampled_from([(5, -3, -4), (1, -1, -1), (2, -2, -2), (3, -3, -1), (1, 0, 0), (0, 0, 0),
                              (2, 1, -1), (5, -3, 0), (-1, -2, -3), (7, -5, -2), (2147483647, -3, -4)])


@settings(max_examples=250, deadline=None)
@given(seqs, seqs, score_sets)
def test_oracle_equals_twin(ref, read, sc):
    r = oracle.align(ref, read, *sc)
    assert (r.score, r.cells, r.sites) == sw_twin.align(ref, read, *sc)
    assert oracle.align(ref, read, *sc, lowmem=True) == r
    assert oracle.score(ref, read, *sc) == (r.score, len(r.cells))


@settings(max_examples=120, deadline=None)
@given(low_complexity, low_complexity, score_sets)
def test_oracle_equals_twin_low_complexity(ref, read, sc):
    r = oracle.align(ref, read, *sc)
    assert (r.score, r.cells, r.sites) == sw_twin.align(ref, read, *sc)


def test_map_ref_reduction_matches_twin():
    rnd = random.Random(3)
    ref = "".join(rnd.choice("ACGT") for _ in range(120))
    reads = ["".join(rnd.choice("ACGT") for _ in range(rnd.randint(5, 30))) for _ in range(9)] + [ref[30:60]]
    total, sites = sw_twin.map_ref(ref, reads)
    out = oracle.cpu_baseline([ref], reads, threads=2, mode=0, want_scores=True)
    assert out["ref_totals"] == [total]
    assert out["pair_scores"] == [sw_twin.align(ref, q)[0] for q in reads]
    # both harness shapes walk the same pairs
    assert oracle.cpu_baseline([ref] * 5, reads, threads=3, mode=1)["checksum"] != 0


@settings(max_examples=150, deadline=None)
@given(seqs, seqs, st.sampled_from([(5, -3, -4), (1, -1, -1), (2, -2, -2), (3, -3, -1), (1, 0, 0), (2, 1, -1)]))
def test_distributed_sw_variant_oracle_equals_twin(ref, read, sc):
    """N3: DistributedSW semantics (strict '>' ties, diagonal-major list, stable sort by beginning)."""
    r = oracle.align(ref, read, *sc, tie_gt=True)
    assert (r.score, r.cells, r.sites) == sw_twin.align_gt(ref, read, *sc)
    assert r.score == oracle.align(ref, read, *sc).score                     # the tie rule never changes H


def test_distributed_sw_known_answers():
    assert oracle.align("GTTCA", "CTA", tie_gt=True).sites == [(3, "TCA", "T_A")]
    assert oracle.align("CCAAT", "CAT", tie_gt=True).sites == [(2, "CAAT", "CA_T")]    # ">=" gives C_AT
    r = oracle.align("ATATATAT", "ATAT", tie_gt=True)
    assert r.score == 20 and [s[0] for s in r.sites] == [1, 3, 5]
