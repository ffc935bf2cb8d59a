"""GPU parity tests: CUDA path through the C ABI vs the CPU oracle, bit-exact
(scores, max-cell lists in list order, beginnings, both alignment strings)."""
import random

import pytest

import oracle
from sparksmithwaterman_b200 import synth
from tests.helpers import check_pairs

pytestmark = pytest.mark.gpu

REF = "CCTGGGTCCTGCCTCGCATCTGACCAGGGCAGGTGGCCTCCTCATCACACTGCTGCCTCTGCTGTTGGCCCTGCTCATGA"
READ_80 = "AATTTTAGTCTCTCCCTACCCTTTTGGACAGAGCTTCCTGTCCTCTCATTTCACAGGTTATGCAACAGAGGGTTCTGTGT"
READ_20 = "ACTGACTGACTGACTGACTG"


def test_kat_small(engine):
    kats = [("ACGT", "ACGT"), ("AAAA", "CCCC"), ("ACGTACGT", ""), ("AC", "ACGTT"), ("ATATATAT", "ATAT"),
            ("acgtTTacgt", "ACGT"), ("GATTACA", "GCATGCU"), ("CGTGAATTCAT", "GACTTAC"), ("AAGGAA", "AAA"),
            ("GTTCA", "CTA"), ("CCAAT", "CAT"), ("ATCAA", "TAC"), ("TGGTC", "TGT")]
    for ref, read in kats:
        check_pairs(engine, [ref], [read])


def test_engineer_data_constants(engine):
    refs = [REF * 5, REF, REF * 2]
    reads = [READ_80, READ_20, READ_20 * 2, READ_20 * 5]
    check_pairs(engine, refs, reads)


def test_tie_heavy(engine):
    refs = ["AT" * 400, "A" * 300, "ACG" * 100, "TA" * 37 + "G" + "AT" * 50]
    reads = ["AT" * 75, "TA" * 75, "A" * 50, "A" * 150, "ACG" * 40, "AT" * 30 + "C" + "AT" * 30]
    check_pairs(engine, refs, reads)


def test_random_small_all_lengths(engine):
    rnd = random.Random(7)
    refs = ["".join(rnd.choice("ACGT") for _ in range(n)) for n in (1, 2, 7, 8, 9, 31, 32, 33, 63, 64, 65, 127, 128, 129, 300, 0)]
    reads = ["".join(rnd.choice("ACGT") for _ in range(m)) for m in (1, 2, 3, 8, 9, 31, 32, 33, 63, 64, 65, 100, 104, 105, 127, 128, 129, 150, 152, 153, 200, 255, 256, 0)]
    check_pairs(engine, refs, reads)


def test_mixed_case_and_foreign_symbols(engine):
    refs = ["acgtACGTacgtTTGACA", "ACGTNNNN".replace("N", "A"), "ggggCCCC"]
    reads = ["ACGTacgt", "nnACGTnn", "GGGGcccc", "TTgaca"]
    check_pairs(engine, refs, reads)


@pytest.mark.parametrize("scores", [(5, -3, -4), (1, -1, -1), (2, -2, -2), (3, -3, -1), (2, -1, -3), (1, 0, -1), (7, -5, -2)])
def test_custom_scores(engine, scores):
    rnd = random.Random(11)
    refs = ["".join(rnd.choice("ACGT") for _ in range(rnd.randint(20, 200))) for _ in range(6)]
    reads = ["".join(rnd.choice("ACGT") for _ in range(rnd.randint(5, 120))) for _ in range(7)]
    reads.append(refs[0][10:60])
    check_pairs(engine, refs, reads, scores)


def test_cfg1_shape(engine):
    """BASELINE config 1: one 100 bp read vs 1,000 RefSeq-shaped refs, every pair, every string."""
    refs, reads = synth.workload(1, 100, 1000)
    n = check_pairs(engine, [r.decode() for r in refs], [q.decode() for q in reads])
    assert n == 1000


def test_cfg2_sample(engine):
    """BASELINE config 2 shape at reduced count: 150 bp reads vs RefSeq-shaped refs."""
    refs, reads = synth.workload(24, 150, 60)
    check_pairs(engine, [r.decode() for r in refs], [q.decode() for q in reads])
