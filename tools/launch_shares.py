#!/usr/bin/env python
"""Per-kernel shares of an `ncu --metrics gpu__time_duration.sum --csv` launch list.
    python tools/launch_shares.py launches.csv [top_n]"""
import collections, csv, sys


def main():
    rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg, cnt = {}, collections.Counter()
    for r in rows[1:]:
        v = float(r[vi].replace(",", ""))
        u = r[ui]
        ms = v / 1e6 if u.startswith("n") else (v / 1e3 if u.startswith("u") else v)
        k = r[ki][:72]
        agg[k] = agg.get(k, 0.0) + ms
        cnt[k] += 1
    tot = sum(agg.values())
    for k, v in sorted(agg.items(), key=lambda x: -x[1])[:top]:
        print(f"{v:10.3f} ms {100 * v / tot:5.1f}% x{cnt[k]:3d} {k}")


if __name__ == "__main__":
    main()
