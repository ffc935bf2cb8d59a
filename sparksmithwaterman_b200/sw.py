"""Host-side mirror of the reference's operator API for the hot path.

The reference is Java; no JVM exists in this image, so this Python module plays the role
of the thin Java host layer (java/sw/SmithWaterman.java): same class and
method names, same argument meaning, same result shape -- the body marshals to the C ABI.

    SmithWaterman.OptAlignments().call(seqs, alignScores, alignTypes)
        reference: /root/reference/src/sw/SmithWaterman.java:35,62-92
        seqs = [reference, read]; alignScores = [match, mismatch, gap];
        alignTypes = [alignment, insertion, deletion, none] (ignored: they never leave the
        operator -- only their identity matters inside GetCellScore / GetAlignment)
        returns (maxScore, [(beginning, [refAligned, readAligned]), ...]) in max-cell order
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

from .engine import Engine

ALIGN_SCORES = [5, -3, -4]              # Distribution.java:36
ALIGN_TYPES = ["a", "i", "d", "-"]      # Distribution.java:37

_engine = None


def default_engine() -> Engine:
    """Process-wide engine on cuda:0 (raises without a GPU: there is no CPU path)."""
    global _engine
    if _engine is None:
        _engine = Engine(0)
    return _engine


def set_default_engine(eng: Engine) -> None:
    global _engine
    _engine = eng


def expand_pair(res, ref_idx: int, read_idx: int) -> Tuple[int, List[Tuple[int, List[str]]]]:
    """One pair of a batched result in the shape OptAlignments.call returns."""
    score, _cells, sites = res.pair(ref_idx, read_idx)
    return score, [(b, [ra, qa]) for (b, ra, qa) in sites]


class SmithWaterman:
    class OptAlignments:
        """Function3<String[], int[], char[], Tuple2<Integer, ArrayList<Tuple2<Integer,String[]>>>>"""

        def call(self, seqs: Sequence[str], alignScores: Sequence[int] = ALIGN_SCORES,
                 alignTypes: Sequence[str] = ALIGN_TYPES):
            if len(seqs) != 2 or len(alignScores) != 3:
                raise ValueError("seqs = [reference, read], alignScores = [match, mismatch, gap]")
            # one pair = one request to the engine's submission queue: calls from concurrent host threads (the
            # unchanged driver's MapRef tasks, Distribution.java:419-426) are coalesced into one launch sequence
            res = default_engine().align_pair(seqs[0], seqs[1], tuple(alignScores)).cache()
            try:
                return expand_pair(res, 0, 0)
            finally:
                res.free()


def distributed_order(cells, sites):
    """DistributedSW's result order: max cells anti-diagonal by anti-diagonal (i + j ascending), j ascending
    inside one (DistributedSW.java:192-245, :907-911), then the alignments stably sorted by beginning
    (:456-490, MatchSiteComp).  Input: row-major cells and their sites."""
    order = sorted(range(len(cells)), key=lambda k: (cells[k][0] + cells[k][1], cells[k][1]))
    order.sort(key=lambda k: sites[k][0])
    return [cells[k] for k in order], [sites[k] for k in order]


class DistributedSW:
    """Drop-in for the reference's other operator, sw.DistributedSW.OptAlignments
    (src/sw/DistributedSW.java:50,77-104): identical signature, strict-'>' tie rule.  The reference
    runs one Spark job per anti-diagonal; here it is the same CUDA path with SWB_F_TIE_GT."""

    class OptAlignments:
        def call(self, seqs: Sequence[str], alignScores: Sequence[int] = ALIGN_SCORES,
                 alignTypes: Sequence[str] = ALIGN_TYPES):
            res = default_engine().align_pair(seqs[0], seqs[1], tuple(alignScores), tie_gt=True).cache()
            try:
                score, cells, sites = res.pair(0, 0)
                _, sites = distributed_order(cells, sites)
                return score, [(b, [ra, qa]) for (b, ra, qa) in sites]
            finally:
                res.free()
