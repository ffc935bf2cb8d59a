"""oracle/sw_twin.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Independent second restatement of the reference's per-pair operator, written in a
deliberately different style from oracle/sw_oracle.c so that the two can check
each other (the Java reference cannot run here):

* the score matrix is filled from the recurrence H = max(0, W+g, N+g, NW+s)
  (README.md:51-92 of the reference; SmithWaterman.java:217-252),
* NO type matrix is stored.  The traceback re-derives the type of a positive cell
  from the scores alone: it is the first of (alignment, insertion, deletion) whose
  candidate equals H -- which is what the ">=" cascade at SmithWaterman.java:228,
  236,245 amounts to (the last passing test wins, tests run d, i, a),
* max cells are collected by a separate row-major scan after the fill
  (equivalent to the clear/append bookkeeping at SmithWaterman.java:176-185).

Pure Python: small inputs only.
"""
from __future__ import annotations


def _wrap32(x: int) -> int:
    x &= 0xFFFFFFFF
    return x - (1 << 32) if x & 0x80000000 else x


def _same(a: str, b: str) -> bool:
    # Character.toUpperCase on ASCII (SmithWaterman.java:311-317)
    return a.upper() == b.upper() if (ord(a) < 128 and ord(b) < 128) else a == b


def fill(ref: str, read: str, match: int, mismatch: int, gap: int):
    n, m = len(ref), len(read)
    H = [[0] * (n + 1) for _ in range(m + 1)]
    for i in range(1, m + 1):
        row, up = H[i], H[i - 1]
        rb = read[i - 1]
        for j in range(1, n + 1):
            s = match if _same(ref[j - 1], rb) else mismatch
            row[j] = max(0, _wrap32(row[j - 1] + gap), _wrap32(up[j] + gap), _wrap32(up[j - 1] + s))
    return H


def align(ref: str, read: str, match: int = 5, mismatch: int = -3, gap: int = -4):
    """Returns (score, [(i, j)], [(beginning, ref_aln, read_aln)])."""
    n, m = len(ref), len(read)
    H = fill(ref, read, match, mismatch, gap)
    best = 0
    for i in range(1, m + 1):
        for j in range(1, n + 1):
            if H[i][j] > best:
                best = H[i][j]
    cells = [(i, j) for i in range(1, m + 1) for j in range(1, n + 1) if H[i][j] == best]
    sites = []
    for (i, j) in cells:
        ra, qa = [], []
        beginning = 0
        while H[i][j] > 0:
            beginning = j
            h = H[i][j]
            s = match if _same(ref[j - 1], read[i - 1]) else mismatch
            if _wrap32(H[i - 1][j - 1] + s) == h:
                ra.append(ref[j - 1]); qa.append(read[i - 1]); i -= 1; j -= 1
            elif _wrap32(H[i - 1][j] + gap) == h:
                ra.append("_"); qa.append(read[i - 1]); i -= 1
            else:
                ra.append(ref[j - 1]); qa.append("_"); j -= 1
        sites.append((beginning, "".join(reversed(ra)), "".join(reversed(qa))))
    return best, cells, sites


def map_ref(ref: str, reads, match: int = 5, mismatch: int = -3, gap: int = -4):
    """Distribution.java:403-436: wrapping int32 total over reads; all sites, stably
    sorted by beginning (comparator Distribution.java:691-694)."""
    total = 0
    sites = []
    for read in reads:
        score, _, s = align(ref, read, match, mismatch, gap)
        total = _wrap32(total + score)
        sites.extend(s)
    sites.sort(key=lambda t: t[0])  # Python's sort is stable, like Collections.sort
    return total, sites


def align_gt(ref: str, read: str, match: int = 5, mismatch: int = -3, gap: int = -4):
    """DistributedSW.OptAlignments (reference src/sw/DistributedSW.java:77-104): same scores; ties between
    candidates go to the FIRST of deletion, insertion, alignment (strict ">" cascade, :300-330); max cells
    listed anti-diagonal by anti-diagonal, j ascending (:192-245, :907-911); alignments stably sorted by
    beginning (:456-490).  Returns (score, cells, sites) in that final order."""
    n, m = len(ref), len(read)
    H = fill(ref, read, match, mismatch, gap)
    best = max([0] + [H[i][j] for i in range(1, m + 1) for j in range(1, n + 1)])
    cells = [(d - j, j) for d in range(2, m + n + 1) for j in range(max(1, d - m), min(n, d - 1) + 1)
             if H[d - j][j] == best] if m and n else []
    out = []
    for (ci, cj) in cells:
        i, j, ra, qa, beginning = ci, cj, [], [], 0
        while H[i][j] > 0:
            beginning = j
            h = H[i][j]
            s = match if _same(ref[j - 1], read[i - 1]) else mismatch
            if _wrap32(H[i][j - 1] + gap) == h:
                ra.append(ref[j - 1]); qa.append("_"); j -= 1
            elif _wrap32(H[i - 1][j] + gap) == h:
                ra.append("_"); qa.append(read[i - 1]); i -= 1
            else:
                assert _wrap32(H[i - 1][j - 1] + s) == h
                ra.append(ref[j - 1]); qa.append(read[i - 1]); i -= 1; j -= 1
        out.append(((ci, cj), (beginning, "".join(reversed(ra)), "".join(reversed(qa)))))
    out.sort(key=lambda t: t[1][0])
    return best, [c for c, _ in out], [s for _, s in out]
