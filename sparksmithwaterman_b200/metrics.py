"""Host-side mirror of the reference's data-set reporter (second half):
metrics.RunningMedian (reference src/metrics/RunningMedian.java:111-171) and metrics.RefSetInfo
(src/metrics/RefSetInfo.java:56-116 getInfo, :129-166 printAllInfo, :178-196 the tables).

Same observable behaviour, quirks included: the running median routes a value by comparing it with the
CURRENT median (`value < median` goes to the lower heap, :114), starts from median 0, and is the mean
of the two middle values for an even count; the report text uses Java's `%-,11d` / `%-,7.2f` layouts.
The sequences come from inout.get_ref_seqs (InOutOps.GetRefSeqs), so a directory summarised here is
the directory the aligner would load.
"""
from __future__ import annotations

import heapq
import os
from typing import Iterator, List, Sequence, Tuple

from . import inout

REF_DIR = "/home/ubuntu/project/reference"          # RefSetInfo.java:32
DELIMITER = ">gi"                                   # RefSetInfo.java:33


class RunningMedian:
    """Two heaps: `_max` holds the values below the median (negated), `_min` the rest."""

    def __init__(self):
        self._max: List[int] = []
        self._min: List[int] = []
        self._median = 0.0

    def add(self, value: int) -> None:
        if value < self._median:                     # RunningMedian.java:114
            heapq.heappush(self._max, -value)
        else:
            heapq.heappush(self._min, value)
        # balance: the sizes differ by at most one (:131-145)
        if len(self._max) > len(self._min) + 1:
            heapq.heappush(self._min, -heapq.heappop(self._max))
        elif len(self._min) > len(self._max) + 1:
            heapq.heappush(self._max, -heapq.heappop(self._min))
        # median (:151-171)
        if len(self._max) == len(self._min):
            self._median = (-self._max[0] + self._min[0]) / 2.0
        elif len(self._max) > len(self._min):
            self._median = float(-self._max[0])
        else:
            self._median = float(self._min[0])

    def get_running_median(self) -> float:
        return self._median


def crawl(root: str) -> Iterator[str]:
    """Every regular file below root, depth first, a directory's entries in listing order
    (sw.DirectoryCrawler; java.io.File.listFiles order is unspecified, os.listdir's too)."""
    if not os.path.exists(root):
        raise FileNotFoundError("Root directory not found: " + root)      # the reference prints this and exits
    for name in os.listdir(root):
        path = os.path.join(root, name)
        if os.path.isdir(path):
            yield from crawl(path)
        else:
            yield path


def info_of_lengths(lengths: Sequence[int]) -> Tuple[List[int], List[float]]:
    """([count, total, min, max], [mean, median]) of a list of sequence lengths, as getInfo accumulates them
    (RefSetInfo.java:92-112); an empty list keeps the reference's initial values (min = Long.MAX_VALUE)."""
    rm = RunningMedian()
    total, lo, hi = 0, (1 << 63) - 1, 0
    for bp in lengths:
        total += bp
        rm.add(bp)
        lo = min(lo, bp)
        hi = max(hi, bp)
    mean = total / len(lengths) if lengths else float("nan")             # Java: 0.0 / 0 = NaN
    return [len(lengths), total, lo, hi], [mean, rm.get_running_median()]


class RefSetInfo:
    @staticmethod
    def get_info(directory: str | None):
        """(directory, n_files, [n_seqs, total_bp, min_bp, max_bp], [mean, median], [(file name, n_seqs)])"""
        if directory is None or not directory.strip():
            directory = REF_DIR
        lengths: List[int] = []
        files: List[Tuple[str, int]] = []
        for path in crawl(directory):
            refs = inout.get_ref_seqs(path, DELIMITER)
            files.append((path.split("/")[-1], len(refs)))
            lengths.extend(len(r[1]) for r in refs)
        longs, doubles = info_of_lengths(lengths)
        return directory, len(files), longs, doubles, files

    @staticmethod
    def formatted_table(table: Sequence[Tuple[str, int]]) -> str:
        out = ["%-35s%1s%11s" % ("File Name", "|", "# Sequences"), "-----------------------------------+-----------"]
        out += ["%-35s%1s%11s" % (name, "|", f"{n:,}") for name, n in table]
        return "\n".join(out) + "\n"

    @staticmethod
    def all_info_str(directory: str | None) -> str:
        d, n_files, longs, doubles, files = RefSetInfo.get_info(directory)
        s = [f"directory = {d}", "", f"# files  =  {n_files}",
             "%-21s  %1s  %-11s" % ("# reference sequences", "=", f"{longs[0]:,}"),
             "%-21s  %1s  %-11s" % ("# total base pairs", "=", f"{longs[1]:,}"),
             "", "base pairs in a sequence:", "-------------------------",
             "%-6s  %1s  %-10s" % ("min", "=", f"{longs[2]:,}"),
             "%-6s  %1s  %-10s" % ("max", "=", f"{longs[3]:,}"),
             "%-6s  %1s  %-7s" % ("mean", "=", f"{doubles[0]:,.2f}"),
             "%-6s  %1s  %-7s" % ("median", "=", f"{doubles[1]:,.2f}"), "", ""]
        text = "\n".join(s) + "\n"
        by_name = sorted(files, key=lambda t: t[0])                       # FilenameComparator (String.compareTo on ASCII names)
        text += RefSetInfo.formatted_table(by_name)
        text += "\n\n"
        by_size = sorted(by_name, key=lambda t: t[1])                     # NumRefComparator on the list sorted above (stable)
        text += RefSetInfo.formatted_table(by_size)
        return text

    @staticmethod
    def print_all_info(directory: str | None, output_file: str) -> None:
        with open(output_file, "w") as f:
            f.write(RefSetInfo.all_info_str(directory))
