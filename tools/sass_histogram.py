#!/usr/bin/env python
"""Per-kernel SASS opcode histogram of libswb200.so (cuobjdump -sass), so that the instruction mix behind the
roofline numbers (DPX VIADDMNMX / VIMNMX3, IMAD on the FMA pipe, LDS.128 / STG.128, SHFL) is committed evidence.

    python tools/sass_histogram.py [--top 14] [--out profiles/sass_histogram_r02.txt]
"""
import argparse
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "sparksmithwaterman_b200", "_lib", "libswb200.so")


def demangle(names):
    try:
        out = subprocess.run(["cu++filt"] + names, capture_output=True, text=True, check=True).stdout.splitlines()
        return dict(zip(names, out))
    except Exception:
        return {n: n for n in names}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--top", type=int, default=14)
    ap.add_argument("--out", default=None)
    ap.add_argument("--lib", default=LIB)
    a = ap.parse_args()
    sass = subprocess.run(["cuobjdump", "-sass", a.lib], capture_output=True, text=True, check=True).stdout
    hist, arch, cur = collections.OrderedDict(), {}, None
    cur_arch = "?"
    for line in sass.splitlines():
        m = re.match(r"\s*arch = (sm_\w+)", line)
        if m:
            cur_arch = m.group(1)
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            hist[cur] = collections.Counter()
            arch[cur] = cur_arch
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Za-z0-9_.]+)", line)
        if m and cur:
            hist[cur][m.group(1)] += 1
    names = demangle(list(hist))
    lines = [f"# cuobjdump -sass {os.path.relpath(a.lib, ROOT)}: static opcode counts per kernel (top {a.top}); "
             f"{len(hist)} kernels, arch {sorted(set(arch.values()))}"]
    total = collections.Counter()
    for k, h in hist.items():
        total.update(h)
        n = sum(h.values())
        short = names[k].replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
        cut = short.rfind(">(") + 1 if ">(" in short else short.find("(")
        short = short[:cut] if cut > 0 else short
        short = short if len(short) <= 110 else short[:107] + "..."
        lines.append(f"\n{short}  [{arch[k]}, {n} instructions]")
        lines.append("  " + "  ".join(f"{op} {c}" for op, c in h.most_common(a.top)))
    lines.append("\n# whole library")
    lines.append("  " + "  ".join(f"{op} {c}" for op, c in total.most_common(40)))
    fams = {"DPX s16x2 (VIADDMNMX/VIMNMX3/VIMNMX .S16x2)": r"^VI(ADDMNMX|MNMX3|MNMX|ADD)\S*(S16x2|16x2)",
            "DPX s32 (VIADDMNMX/VIMNMX3)": r"^VI(ADDMNMX|MNMX3)(?!\S*S16)",
            "IMAD": r"^IMAD", "LDS.128": r"^LDS\S*\.128", "STG.128": r"^STG\S*\.128", "SHFL": r"^SHFL",
            "tcgen05 / UTC*MMA": r"^UTC", "TMA (UBLKCP/UTMA)": r"^(UBLKCP|UTMA)"}
    lines.append("\n# families")
    for name, pat in fams.items():
        lines.append(f"  {name}: {sum(c for op, c in total.items() if re.search(pat, op))}")
    text = "\n".join(lines) + "\n"
    if a.out:
        with open(a.out, "w") as f:
            f.write(text)
    sys.stdout.write(text)


if __name__ == "__main__":
    main()
