"""On-disk formats of the reference (src/sw/InOutOps.java) -- SURVEY.md section 8(f) row N2.

    get_reads(path, delimiter)        InOutOps.GetReads.call      (:49-89)
    get_ref_seqs(path, delimiter)     InOutOps.GetRefSeqs.call    (:100-169)
    get_output_str(reads, ...)        InOutOps.GetOutputStr.call  (:226-289)
    run_reference_file(...)           one (input file x reference file) unit of DistributeReference.call
                                      (Distribution.java:310-367) through the batched native map

Quirks kept on purpose: read lines are trim()med, reference sequence lines are NOT; every
non-first line of an input file is a read (even an empty one); the delimiter is a prefix test
(IsMetadata, :394-412); a reference file whose first line is not a header is an error in the
reference (NullPointerException) and a ValueError here.
"""
from __future__ import annotations

import time
from typing import List, Sequence, Tuple

from . import distribution, sw

NEWLINE = "\n"
TAB = "\t"
DELIMITER = ">gi"                    # Distribution.java:40


def is_metadata(line: str, delimiter: str) -> bool:
    return len(line) >= len(delimiter) and line[:len(delimiter)] == delimiter


def _java_trim(s: str) -> str:
    """String.trim(): strips code points <= U+0020 on both ends."""
    a, b = 0, len(s)
    while a < b and s[a] <= " ":
        a += 1
    while b > a and s[b - 1] <= " ":
        b -= 1
    return s[a:b]


def _lines(path: str) -> List[str]:
    # java.util.Scanner.nextLine(): splits on \r\n, \n, \r (and a few exotic separators)
    with open(path, "r", encoding="latin-1", newline="") as f:
        text = f.read()
    out = text.replace("\r\n", "\n").replace("\r", "\n").split("\n")
    if out and out[-1] == "":
        out.pop()                     # no phantom line after a final terminator
    return out


def get_reads(path: str, delimiter: str = DELIMITER) -> List[str]:
    lines = _lines(path)
    if not lines:
        raise ValueError("empty input file (Scanner.nextLine would throw)")
    reads = []
    first = _java_trim(lines[0])
    if not is_metadata(first, delimiter):
        reads.append(first)
    reads.extend(_java_trim(x) for x in lines[1:])
    return reads


def get_ref_seqs(path: str, delimiter: str = DELIMITER) -> List[List[str]]:
    seqs: List[List[str]] = []
    ref = None
    parts: List[str] = []
    for line in _lines(path):
        if is_metadata(line, delimiter):
            if ref is not None:
                seqs.append([ref, "".join(parts)])
            ref, parts = line, []
        else:
            if ref is None:
                raise ValueError("reference file does not start with a header line")
            parts.append(line)                       # NOT trimmed (InOutOps.java:148)
    if ref is None:
        raise ValueError("reference file holds no header line")
    seqs.append([ref, "".join(parts)])
    return seqs


def get_output_str(reads: Sequence[str], nums: Tuple[int, int], max_score: int, exec_ms: int, opt) -> str:
    """opt = [([metadata, sequence], [(beginning, [refAligned, readAligned]), ...]), ...]"""
    s = ["Execution Time = %d ms" % exec_ms + NEWLINE, NEWLINE,
         "# Reference Sequences = %d" % nums[0] + NEWLINE, "# Reads = %d" % nums[1] + NEWLINE, NEWLINE,
         "Input:" + NEWLINE]
    s.extend(r + NEWLINE for r in reads)
    s.append(NEWLINE)
    s.append("Maximum alignment score = %d" % max_score)
    s.append(NEWLINE)
    for (ref, sites) in opt:
        s.append("Reference:" + NEWLINE)
        s.append(ref[0] + NEWLINE)
        s.append(ref[1] + NEWLINE)
        s.append(NEWLINE)
        for (beginning, aligned) in sites:
            s.append(TAB + "Index = %d" % beginning + NEWLINE)
            s.append(TAB + aligned[0] + NEWLINE)
            s.append(TAB + aligned[1] + NEWLINE)
            s.append(NEWLINE)
    return "".join(s)


def run_reference_files(ref_paths: Sequence[str], input_path: str, delimiter: str = DELIMITER,
                        align_scores=sw.ALIGN_SCORES, engine=None, as_written: bool = True) -> str:
    """One input file against a list of reference files -> the result text the reference writes.
    as_written=True keeps DistributeReference's first()/lookup() reduce (Distribution.java:341-352);
    False uses NoDistribution's true running max (:601-613)."""
    reads = get_reads(input_path, delimiter)
    t0 = time.perf_counter()
    n_refs = 0
    mapped_per_file = []
    for p in ref_paths:
        refs = get_ref_seqs(p, delimiter)
        n_refs += len(refs)
        mapped_per_file.append(distribution.map_refs(refs, reads, align_scores, engine=engine))
    if as_written:
        best, opt = distribution.DistributeReference.reduce(mapped_per_file)
    else:
        best, opt = distribution.NoDistribution.reduce([m for f in mapped_per_file for m in f])
    ms = int((time.perf_counter() - t0) * 1000)
    return get_output_str(reads, (n_refs, len(reads)), best, ms, opt)
