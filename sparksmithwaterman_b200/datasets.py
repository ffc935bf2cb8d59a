"""The reference's four benchmark sweeps (src/metrics/EngineerData.java:51-224, driven by
ExecutionTimesReference.runTest1..4) as in-memory generators -- SURVEY.md section 8(f) row N4.

The three constants are the reference's only shipped fixtures (EngineerData.java:23,26,29).
The counterpart sets EngineerData does not generate (the references of tests 1-2, the reads of
tests 3-4 live in unshipped directories) are stated assumptions: 1,000 x 400 bp references and
5 x READ_80 reads."""
from __future__ import annotations

from typing import Iterator, List, Tuple

REF = "CCTGGGTCCTGCCTCGCATCTGACCAGGGCAGGTGGCCTCCTCATCACACTGCTGCCTCTGCTGTTGGCCCTGCTCATGA"
READ_80 = "AATTTTAGTCTCTCCCTACCCTTTTGGACAGAGCTTCCTGTCCTCTCATTTCACAGGTTATGCAACAGAGGGTTCTGTGT"
READ_20 = "ACTGACTGACTGACTGACTG"

DEFAULT_REFS = [REF * 5] * 1000
DEFAULT_READS = [READ_80] * 5

Point = Tuple[str, int, List[str], List[str]]      # (sweep, x, refs, reads)


def change_read_num() -> Iterator[Point]:
    """EngineerData.changeReadNum (:51-79): 20 reads, then 50, 100, ..., 1600 reads of READ_80."""
    yield ("read_num", 20, DEFAULT_REFS, [READ_80] * 20)
    n = 0
    while n <= 1574:
        n += 50
        yield ("read_num", n, DEFAULT_REFS, [READ_80] * n)


def change_read_len() -> Iterator[Point]:
    """EngineerData.changeReadLen (:87-104): 5 reads of READ_20 x k, 20 .. 500 bp."""
    length = 0
    while length < 500:
        length += 20
        yield ("read_len", length, DEFAULT_REFS, [READ_20 * (length // 20)] * 5)


def change_ref_num() -> Iterator[Point]:
    """EngineerData.changeRefNum (:116-169): 1, 10, 30, 50, 100, 500, 1000, 1500, 2000, then +2000 up to
    40,000 references of 400 bp (REF x 5)."""
    nums = [1, 10, 30, 50, 100, 500, 1000, 1500, 2000]
    n = 2000
    while n < 40000:
        n += 2000
        nums.append(n)
    for n in nums:
        yield ("ref_num", n, [REF * 5] * n, DEFAULT_READS)


def change_ref_len() -> Iterator[Point]:
    """EngineerData.changeRefLen (:178-224): one reference of 1, 5, 10, 20 lines of REF, then 50, 100,
    ..., 1600 lines (80 bp .. 128 kbp)."""
    for rows in (1, 5, 10, 20):
        yield ("ref_len", rows * 80, [REF * rows], DEFAULT_READS)
    lines = 0
    while lines <= 1574:
        lines += 50
        yield ("ref_len", lines * 80, [REF * lines], DEFAULT_READS)


SWEEPS = {"read_num": change_read_num, "read_len": change_read_len, "ref_num": change_ref_num,
          "ref_len": change_ref_len}
