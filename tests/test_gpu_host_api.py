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


def test_reference_file_round_trip(engine, tmp_path):
    """N2: reference-format files in, the reference's result text out (both reductions)."""
    from sparksmithwaterman_b200 import inout
    rnd = random.Random(13)
    refs = [[f">gi|{k}|syn|", "".join(rnd.choice("ACGT") for _ in range(rnd.randint(60, 300)))] for k in range(5)]
    reads = [refs[2][1][10:50], "".join(rnd.choice("ACGT") for _ in range(30)), refs[4][1][0:25]]
    rp = tmp_path / "ref1.fa"
    rp.write_text("".join(f"{m}\n{s[:40]}\n{s[40:]}\n" for m, s in refs))
    ip = tmp_path / "in1.txt"
    ip.write_text(">gi reads\n" + "\n".join(reads) + "\n")
    assert inout.get_ref_seqs(str(rp)) == refs and inout.get_reads(str(ip)) == reads
    mapped = []
    for meta, seq in refs:
        total, sites = sw_twin.map_ref(seq, reads)
        mapped.append((total, ([meta, seq], [(b, [ra, qa]) for (b, ra, qa) in sites])))
    for as_written, reduce in ((True, lambda m: distribution.DistributeReference.reduce([m])),
                               (False, distribution.NoDistribution.reduce)):
        txt = inout.run_reference_files([str(rp)], str(ip), engine=engine, as_written=as_written)
        best, opt = reduce(mapped)
        exp = inout.get_output_str(reads, (5, 3), best, 0, opt)
        strip = lambda t: t.split("\n", 1)[1]                    # drop the execution-time line
        assert strip(txt) == strip(exp)


def test_engineer_data_sweep_points(engine):
    """N4: a few points of the reference's own sweeps, bit-exact."""
    from sparksmithwaterman_b200 import datasets
    from tests.helpers import check_pairs
    pts = [next(datasets.change_read_len()), next(datasets.change_ref_len())]
    it = datasets.change_read_len()
    for _ in range(12):
        p = next(it)
    pts.append(p)                                                # 240 bp reads
    for sweep, x, refs, reads in pts:
        check_pairs(engine, refs[:3], reads[:2])
    assert [p[1] for p in datasets.change_ref_num()][:10] == [1, 10, 30, 50, 100, 500, 1000, 1500, 2000, 4000]
    assert len(list(datasets.change_ref_num())) == 28 and len(list(datasets.change_ref_len())) == 36   # runTest3/4 loop counts
    assert len(list(datasets.change_read_num())) == 33 and len(list(datasets.change_read_len())) == 25


def test_distributed_sw_tie_mode(engine):
    """N3: DistributedSW.OptAlignments (strict '>' ties, diagonal-major order, stable sort by beginning)."""
    sw.set_default_engine(engine)
    op = sw.DistributedSW.OptAlignments()
    assert op.call(["GTTCA", "CTA"]) == (6, [(3, ["TCA", "T_A"])])            # '>=' gives (4, C_A, CTA)
    assert sw.SmithWaterman.OptAlignments().call(["GTTCA", "CTA"]) == (6, [(4, ["C_A", "CTA"])])
    rnd = random.Random(17)
    cases = [("ATATATAT", "ATAT"), ("AAGGAA", "AAA"), ("CGTGAATTCAT", "GACTTAC"), ("AT" * 60, "AT" * 20)]
    cases += [("".join(rnd.choice("ACGT") for _ in range(rnd.randint(10, 300))),
               "".join(rnd.choice("ACGT") for _ in range(rnd.randint(5, 300)))) for _ in range(25)]
    for scores in ((5, -3, -4), (1, -1, -1), (2, -1, -2)):
        for ref, read in cases:
            exp = oracle.align(ref, read, *scores, tie_gt=True)
            got = op.call([ref, read], list(scores))
            assert got == (exp.score, [(b, [ra, qa]) for (b, ra, qa) in exp.sites]), (ref[:20], read[:20], scores)


def test_scores_only_mode_is_exact(engine):
    """SWB_F_SCORES_ONLY: no max-cell lists, no traceback -- but the scores must be the exact maxima.  With the
    subsampled tile maxima of the default fill that needs the locate scan pass; other score sets take the exact fill."""
    import numpy as np
    rnd = random.Random(23)
    base = "".join(rnd.choice("ACGT") for _ in range(3000))
    refs = [base[:1200], base[1000:2500], "".join(rnd.choice("ACGT") for _ in range(700)), "AT" * 150, base[::-1][:900]]
    reads = [base[100:250], base[1100:1181], base[2000:2140][::-1], "AT" * 40, base[500:533], base[1500:1800],
             "".join(rnd.choice("ACGT") for _ in range(150))]
    rs = engine.load_refset(refs)
    for scores in ((5, -3, -4), (3, 1, -2), (2, -2, -2), (5, -9, -4)):
        so = rs.align(reads, scores, scores_only=True)
        full = rs.align(reads, scores)
        exp = np.array([[oracle.align(r, q, *scores, max_cells=1).score for q in reads] for r in refs], dtype=np.int32)
        assert (so.scores == exp).all(), scores
        assert (full.scores == exp).all(), scores
        assert (so.ref_totals == full.ref_totals).all()
        assert (so.best_hits[:, :2] == full.best_hits[:, :2]).all()
        so.free(); full.free()
    rs.free()


def test_submission_queue_coalesces_concurrent_pair_calls(engine):
    """The unchanged driver's path: N host threads each call OptAlignments.call on their own pairs
    (Distribution.java:419-426).  swb_align_pair coalesces them; every caller still gets exactly its pair's
    result (score, cells, beginnings, strings), including score-0 pairs, tie-heavy pairs, mixed case, a
    second score set in the same queue, and a request the engine refuses (only THAT caller sees the error)."""
    from concurrent.futures import ThreadPoolExecutor
    from sparksmithwaterman_b200._ffi import SwbError
    rnd = random.Random(11)
    refs = ["".join(rnd.choice("ACGT") for _ in range(rnd.randint(60, 1500))) for _ in range(24)]
    refs += ["AT" * 90, "acgtTTacgt" * 9, "AAAA", REF * 5]
    reads = ["".join(rnd.choice("ACGT") for _ in range(rnd.randint(20, 150))) for _ in range(10)]
    reads += [refs[3][10:130], "AT" * 30, "CCCC", READ_20, "", refs[1][100:360]]          # the last one: 260 rows, int32 path
    jobs = [(r, q, (5, -3, -4) if (r + q) % 5 else (2, -1, -2)) for r in range(len(refs)) for q in range(len(reads))]
    rnd.shuffle(jobs)
    jobs = jobs[:320] + [("BAD", 0, (5, -3, -4))] * 3
    before = engine.queue_stats()

    def one(job):
        r, q, sc = job
        if r == "BAD":
            try:
                engine.align_pair("AC\xe9GT", reads[q], sc)
            except SwbError as e:
                return "non-ASCII" in str(e)
            return False
        res = engine.align_pair(refs[r], reads[q], sc).cache()
        got = res.pair(0, 0, max_cells=200)
        cnt = res.pair_cell_count(0, 0)
        best = res.best_hits[0].tolist()
        batch = int(res.stats["batches"])
        res.free()
        exp = oracle.align(refs[r], reads[q], *sc, max_cells=200)
        ok = got[0] == exp.score and got[1] == exp.cells and got[2] == exp.sites
        ok &= best[0] == exp.score and (best[2:] == list(exp.cells[0]) if exp.score > 0 else best[2:] == [0, 0])
        if exp.score == 0:
            ok &= cnt == len(refs[r]) * len(reads[q])
        return ok and batch >= 1

    with ThreadPoolExecutor(16) as ex:
        results = list(ex.map(one, jobs))
    assert all(results), [j for j, ok in zip(jobs, results) if not ok][:5]
    after = engine.queue_stats()
    assert after["calls"] - before["calls"] == len(jobs)
    assert after["batches"] - before["batches"] < len(jobs) and after["largest_batch"] >= 2      # coalescing happened
