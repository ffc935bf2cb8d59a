"""Shared parity helper: the CUDA path (through the C ABI) against the C oracle."""
import oracle


def check_pairs(eng, refs, reads, scores=(5, -3, -4), pairs=None, max_cells=None):
    rs = eng.load_refset(refs)
    res = rs.align(reads, scores).cache()
    n_checked = 0
    it = pairs if pairs is not None else [(r, q) for r in range(len(refs)) for q in range(len(reads))]
    for (r, q) in it:
        exp = oracle.align(refs[r], reads[q], *scores, max_cells=max_cells)
        got = res.pair(r, q, max_cells=max_cells)
        assert got[0] == exp.score, f"score ref={r} read={q}: {got[0]} != {exp.score}"
        exp_n = len(refs[r]) * len(reads[q]) if exp.score == 0 else None
        if exp_n is not None:
            assert res.pair_cell_count(r, q) == exp_n
        elif max_cells is None:
            assert res.pair_cell_count(r, q) == len(exp.cells)
        assert got[1] == exp.cells, f"cells ref={r} read={q}: {got[1][:5]} != {exp.cells[:5]}"
        assert got[2] == exp.sites, f"sites ref={r} read={q}: {got[2][:2]} != {exp.sites[:2]}"
        n_checked += 1
    # per-ref wrapping totals (Distribution.java:424)
    import numpy as np
    sc = res.scores.astype(np.int64)
    tot = (sc.sum(axis=1) & 0xFFFFFFFF).astype(np.uint32).view(np.int32) if sc.size else np.zeros(len(refs), np.int32)
    assert (res.ref_totals == tot).all()
    res.free(); rs.free()
    return n_checked
