"""GPU: the host-side mirror of the reference's operator API (sw.SmithWaterman.OptAlignments,
Distribution.MapRef / map_refs, the two reductions) through the C ABI, against the oracle."""
import random

import pytest

import oracle
from oracle import sw_twin
from sparksmithwaterman_b200 import distribution, sw

pytestmark = pytest.mark.gpu

REF = "CCTGGGTCCTGCCTCGCATCTGACCAGGGCAGGTGGCCTCCTCATCACACTGCTGCCTCTGCTGTTGGCCCTGCTCATGA"
READ_20 = "ACTGACTGACTGACTGACTG"


def test_opt_alignments_call(engine):
    sw.set_default_engine(engine)
    op = sw.SmithWaterman.OptAlignments()
    score, opt = op.call([REF * 5, READ_20], [5, -3, -4], ["a", "i", "d", "-"])
    exp = oracle.align(REF * 5, READ_20)
    assert score == exp.score == 56
    assert opt == [(b, [ra, qa]) for (b, ra, qa) in exp.sites]
    assert opt[0] == (49, ["ACTG_CTGCCT__CTG_CTG", "ACTGACTGACTGACTGACTG"])
    # score 0: every cell, empty alignments (SmithWaterman.java:180-185)
    score, opt = op.call(["AAAA", "CCCC"], [5, -3, -4], ["a", "i", "d", "-"])
    assert score == 0 and opt == [(0, ["", ""])] * 16
    assert op.call(["ACGT", ""], [5, -3, -4], ["a", "i", "d", "-"]) == (0, [])


def test_map_refs_matches_mapref_semantics(engine):
    rnd = random.Random(5)
    refs = [[f">gi|{k}|syn|", "".join(rnd.choice("ACGT") for _ in range(rnd.randint(40, 400)))] for k in range(7)]
    refs.append([">gi|rep|", "AT" * 60])
    reads = ["".join(rnd.choice("ACGT") for _ in range(rnd.randint(10, 90))) for _ in range(9)]
    reads += [refs[2][1][5:70], "AT" * 20, "TA" * 20]
    mapped = distribution.map_refs(refs, reads, engine=engine)
    for (total, (ref, sites)), r in zip(mapped, refs):
        assert ref == r
        exp_total, exp_sites = sw_twin.map_ref(r[1], reads)
        assert total == exp_total
        assert sites == [(b, [ra, qa]) for (b, ra, qa) in exp_sites]       # stable sort by beginning
    best, opt = distribution.NoDistribution.reduce(mapped)
    assert best == max(t for t, _ in mapped)
    assert [v[0][0] for v in opt] == sorted(v[0][0] for t, v in mapped if t == best)
    # MapRef.call on one tuple == the batched map for that ref
    sw.set_default_engine(engine)
    one = distribution.MapRef().call((refs[3], reads, ([5, -3, -4], ["a", "i", "d", "-"])))
    assert one == mapped[3]


def test_best_hits_and_totals(engine):
    rnd = random.Random(9)
    refs = ["".join(rnd.choice("ACGT") for _ in range(rnd.randint(60, 500))) for _ in range(25)]
    reads = [refs[k][10:10 + 40] for k in (3, 7, 7, 20)] + ["".join(rnd.choice("ACGT") for _ in range(50)) for _ in range(5)]
    rs = engine.load_refset(refs)
    res = rs.align(reads)
    best = res.best_hits
    for q, read in enumerate(reads):
        scores = [oracle.align(r, read).score for r in refs]
        s = max(scores); k = scores.index(s)                      # lowest ref index on ties
        cell = oracle.align(refs[k], read).cells[0]
        assert tuple(best[q]) == (s, k, cell[0], cell[1])
    res.free(); rs.free()


def test_unsupported_inputs_are_loud(engine):
    from sparksmithwaterman_b200 import SwbError
    rs = engine.load_refset(["ACGT" * 10])
    with pytest.raises(SwbError):
        rs.align(["AC\xe9T"])                                     # non-ASCII byte
    with pytest.raises(SwbError):
        rs.align(["ACGT"], (2**30, -3, -4))                       # could leave int32 (Java would wrap)
    with pytest.raises(SwbError):
        engine.load_refset(["AC\xffT"])
    rs.free()
