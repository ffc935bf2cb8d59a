"""Host-side mirror of the callers of the hot path (reference src/sw/Distribution.java).

    MapRef.call(tuple)            Distribution.java:403-436   one ref x all reads
    map_refs(refs, reads, ...)    the batched form of the same map (one native call for a
                                  whole reference file instead of one per pair)
    NoDistribution.reduce(...)    Distribution.java:584-613   true running max with ties
    DistributeReference.reduce    Distribution.java:341-352   the driver's first()/lookup() reduce
    opt_seqs_sort                 Distribution.java:662-665   sort by metadata
Pure host logic lives in reduce_* / sort helpers so it is testable without a GPU.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

from . import sw

Site = Tuple[int, List[str]]
MapRefOut = Tuple[int, Tuple[List[str], List[Site]]]


def wrap32(x: int) -> int:
    x &= 0xFFFFFFFF
    return x - (1 << 32) if x & 0x80000000 else x


def match_site_sort(sites: List[Site]) -> List[Site]:
    """Collections.sort(matchSites, MatchSiteComp): stable, ascending beginning
    (Distribution.java:428, :691-694)."""
    return sorted(sites, key=lambda t: t[0])


def collect_ref(res, ref_idx: int, n_reads: int) -> Tuple[int, List[Site]]:
    """totalScore and matchSites of one reference from a batched result
    (Distribution.java:419-428)."""
    total = 0
    sites: List[Site] = []
    for q in range(n_reads):
        score, s = sw.expand_pair(res, ref_idx, q)
        total = wrap32(total + score)
        sites.extend(s)
    return total, match_site_sort(sites)


def map_refs(refs: Sequence[Sequence[str]], reads: Sequence[str], align_scores=sw.ALIGN_SCORES,
             engine=None) -> List[MapRefOut]:
    """listRDD.mapToPair(new MapRef()) for a whole list of [metadata, sequence] refs:
    one refset upload + one align call (Distribution.java:337-338)."""
    eng = engine or sw.default_engine()
    rs = eng.load_refset([r[1] for r in refs])
    try:
        res = rs.align(list(reads), tuple(align_scores)).cache()
        try:
            totals = res.ref_totals
            out = []
            for k, ref in enumerate(refs):
                total, sites = collect_ref(res, k, len(reads))
                assert total == int(totals[k])
                out.append((total, (list(ref), sites)))
            return out
        finally:
            res.free()
    finally:
        rs.free()


class MapRef:
    """PairFunction<Tuple3<String[], ArrayList<String>, Tuple2<int[],char[]>>, Integer, ...>"""

    def call(self, tup) -> MapRefOut:
        ref, reads, (align_scores, _align_types) = tup
        return map_refs([ref], reads, align_scores)[0]


class NoDistribution:
    @staticmethod
    def reduce(mapped: Sequence[MapRefOut]):
        """Running max over refs, ties kept in encounter order, then sorted by metadata
        (Distribution.java:601-613, :621)."""
        best, opt = 0, []
        for total, value in mapped:
            if total > best:
                best, opt = total, [value]
            elif total == best:
                opt.append(value)
        return best, opt_seqs_sort(opt)


class DistributeReference:
    @staticmethod
    def reduce(mapped_per_file: Sequence[Sequence[MapRefOut]]):
        """The driver's reduce AS WRITTEN (Distribution.java:341-352): per reference file the
        "max" key is the key of the FIRST mapped element (sortByKey's result is dropped), and
        lookup(maxKey) returns every element of that file with that key."""
        best, opt = 0, []
        for mapped in mapped_per_file:
            if not mapped:
                continue
            max_key = mapped[0][0]
            hits = [v for (k, v) in mapped if k == max_key]
            if max_key > best:
                best, opt = max_key, list(hits)
            elif max_key == best:
                opt.extend(hits)
        return best, opt_seqs_sort(opt)


def opt_seqs_sort(opt):
    """Collections.sort(opt, OptSeqsComp): ascending metadata string (Distribution.java:662-665)."""
    return sorted(opt, key=lambda v: v[0][0])
