"""GPU: BASELINE configs at (reduced but structurally complete) size inside pytest, so the driver's `-m gpu` run
carries them: cfg3 at 20 kbp against the oracle's committed golden results, cfg5 (tie-heavy) invariance over
1/2/4/8 GPUs through the multi-GPU C ABI (swb_multi_*), skipped when the box has fewer devices."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cfg3_20kbp_against_golden(engine):
    """4 pairs of 20 kbp x 20 kbp (3 homologous, 1 random): intra-pair band pipeline (20 bands) + corridor
    traceback; score, max cells, beginnings and both alignment strings equal the oracle's (tests/golden/cfg3_20k.json)."""
    from tests.golden.make_cfg3_golden import cfg3_sequences, site_digest
    with open(os.path.join(ROOT, "tests", "golden", "cfg3_20k.json")) as f:
        golden = json.load(f)
    read, refs = cfg3_sequences(golden["length"], golden["pairs"])
    rs = engine.load_refset(refs)
    res = rs.align([read]).cache()
    for k, g in enumerate(golden["per_pair"]):
        got = res.pair(k, 0)
        assert got[0] == g["score"]
        assert [list(c) for c in got[1]] == g["cells"]
        assert site_digest(got[2]) == g["sites"]
    res.free(); rs.free()


def _device_count():
    import sparksmithwaterman_b200 as swb
    from sparksmithwaterman_b200 import _ffi
    return int(_ffi.load().swb_device_count())


def _cfg5_workload():
    import random
    rnd = random.Random(5)
    refs = []
    for k in range(24):
        unit = ["AT", "A", "ACG", "TA"][k % 4]
        n = int(min(3000, max(40, rnd.lognormvariate(6.3, 0.6))))
        refs.append((unit * (n // len(unit) + 1))[:n])
    refs += ["".join(rnd.choice("ACGT") for _ in range(rnd.randint(100, 900))) for _ in range(8)]
    reads = ["AT" * 75, "TA" * 75, "A" * 150, "AT" * 37 + "C" + "AT" * 37, "ACG" * 50, "CCCC", "AT" * 150,
             "".join(rnd.choice("ACGT") for _ in range(150))]
    return refs, reads


@pytest.mark.parametrize("n_gpus", [1, 2, 4, 8])
def test_cfg5_tie_heavy_identical_across_gpu_counts(engine, n_gpus):
    """cfg5: low-complexity repeats -> thousands of max cells per pair.  Every pair's (score, cells, alignments)
    and the merged best hits from an n-GPU swb_multi_align equal the single-context answer (which the other
    GPU tests pin to the oracle)."""
    if _device_count() < n_gpus:
        pytest.skip(f"needs {n_gpus} GPUs")
    from sparksmithwaterman_b200 import multigpu
    refs, reads = _cfg5_workload()
    rs = engine.load_refset(refs)
    one = rs.align(reads).cache()
    best1 = one.best_hits
    me = multigpu.MultiEngine(list(range(n_gpus)))
    me.load_refset(refs)
    seen = set()
    for d in range(n_gpus):
        seen.update(int(g) for g in me.shard_refs(d))
    assert seen == set(range(len(refs)))
    mr = me.align(reads)
    assert (mr.best_hits == best1).all()
    for r in range(len(refs)):
        for q in range(len(reads)):
            exp = one.pair(r, q, max_cells=300)
            got = mr.pair(r, q, max_cells=300)
            assert got == exp, (r, q)
            sh, loc = me.ref_location(r)
            assert mr.shard(sh).pair_cell_count(loc, q) == one.pair_cell_count(r, q)
    # per-reference totals stay with the owning shard
    tot1 = one.ref_totals
    for d in range(n_gpus):
        ids = me.shard_refs(d)
        assert (mr.shard(d).ref_totals == tot1[ids]).all()
    mr.free(); me.close(); one.free(); rs.free()


def test_multi_more_devices_than_refs_and_score_zero(engine):
    """A shard without references reports (0, -1, 0, 0) and must lose the merge; a read that matches nothing
    keeps the single-GPU answer (score 0, ref 0)."""
    if _device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from sparksmithwaterman_b200 import multigpu
    refs, reads = ["ACGTACGTAC"], ["ACGT", "GGGG", ""]
    rs = engine.load_refset(refs)
    one = rs.align(reads)
    me = multigpu.MultiEngine([0, 1])
    me.load_refset(refs)
    mr = me.align(reads)
    assert (mr.best_hits == one.best_hits).all()
    mr.free(); me.close(); one.free(); rs.free()
