"""GPU: seeded random workloads through the C ABI vs the oracle -- mixed lengths (empty, 1, class
boundaries, > 256 -> wide path, > 2 segments -> segmented fill), alphabets (2, 4, 6 symbols), planted
and low-complexity sequences, score sets on both fill kernels and on the wide path, both tie rules."""
import random
import threading

import pytest

import oracle
from tests.helpers import check_pairs

pytestmark = pytest.mark.gpu

SCORE_SETS = [(5, -3, -4), (5, -3, -4), (1, -1, -1), (2, -2, -2), (3, -3, -1), (5, -9, -4), (2, -1, -3),
              (7, -5, -2), (10, -2, -7), (1, -4, -1), (3, 1, -2), (4, -4, 0), (1, 0, 0)]


def _seq(rnd, n, alphabet):
    mode = rnd.random()
    if mode < 0.15 and n >= 4:                                  # tandem repeat
        unit = "".join(rnd.choice(alphabet) for _ in range(rnd.randint(1, 4)))
        return (unit * (n // len(unit) + 1))[:n]
    return "".join(rnd.choice(alphabet) for _ in range(n))


def _workload(rnd):
    alphabet = rnd.choice(["ACGT", "ACGT", "ACGT", "AT", "ACGTNR", "acgtACGT"])
    ref_lens = [rnd.choice([0, 1, 2, 15, 16, 17, 31, 33, 100, 400, 1023, 1025, 2047, 2049, 2500, 5000])
                for _ in range(rnd.randint(1, 9))]
    if rnd.random() < 0.3:
        ref_lens.append(rnd.randint(6000, 20000))
    refs = [_seq(rnd, n, alphabet) for n in ref_lens]
    read_lens = [rnd.choice([0, 1, 3, 8, 31, 32, 33, 36, 40, 41, 50, 56, 57, 64, 76, 80, 81, 100, 104, 105, 128, 150, 152, 153, 200, 255, 256])
                 for _ in range(rnd.randint(1, 8))]
    if rnd.random() < 0.25:
        read_lens.append(rnd.choice([257, 300, 513, 700]))
    reads = [_seq(rnd, m, alphabet) for m in read_lens]
    for _ in range(rnd.randint(0, 3)):                          # planted, mutated substrings
        r = rnd.choice(refs)
        if len(r) > 40:
            a = rnd.randrange(0, len(r) - 30)
            s = list(r[a:a + rnd.randint(20, min(250, len(r) - a))])
            for _ in range(len(s) // 20):
                p = rnd.randrange(len(s)); s[p] = rnd.choice(alphabet)
            if len(s) > 10 and rnd.random() < 0.5:
                del s[rnd.randrange(len(s))]
            reads.append("".join(s))
    return refs, reads


@pytest.mark.parametrize("seed", range(60))
def test_fuzz_against_oracle(engine, seed):
    rnd = random.Random(1000 + seed)
    refs, reads = _workload(rnd)
    scores = SCORE_SETS[seed % len(SCORE_SETS)]
    check_pairs(engine, refs, reads, scores, max_cells=300)


def test_concurrent_callers_share_one_context(engine):
    """Spark local[N] task threads call the operator concurrently: the context
    serialises them; every thread must get its own exact result."""
    rnd = random.Random(5)
    refs = ["".join(rnd.choice("ACGT") for _ in range(rnd.randint(100, 1200))) for _ in range(12)]
    rs = engine.load_refset(refs)
    errors = []

    def worker(k):
        try:
            r2 = random.Random(k)
            reads = ["".join(r2.choice("ACGT") for _ in range(r2.randint(20, 200))) for _ in range(5)]
            res = rs.align(reads).cache()
            for r in range(len(refs)):
                for q in range(len(reads)):
                    exp = oracle.align(refs[r], reads[q])
                    got = res.pair(r, q)
                    assert got[0] == exp.score and got[1] == exp.cells and got[2] == exp.sites
            res.free()
        except Exception as e:                                   # noqa: BLE001
            errors.append(repr(e))

    threads = [threading.Thread(target=worker, args=(k,)) for k in range(6)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    rs.free()
    assert not errors, errors[:2]
