"""GPU: the fallback kernels on the DEFAULT scores.  The engine picks its kernels from the score set (biased
fill + subsampled tile maxima + byte-tile traceback for 5/-3/-4); environment switches force the other
implementations -- unbiased fill (swb_fill.cu), exact tile maxima, group traceback (swb_trace.cu) -- so that every
kernel is checked on the same workload.  The switches are read once per process: each case runs in a subprocess."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r"""
import random, sys
sys.path.insert(0, %r)
import sparksmithwaterman_b200 as swb
from tests.helpers import check_pairs
rnd = random.Random(77)
base = "".join(rnd.choice("ACGT") for _ in range(2600))
refs = [base, base[300:1500], "".join(rnd.choice("ACGT") for _ in range(900)), "AT" * 220, base[::-1][:1300], "ACG" * 150]
reads = [base[100:250], base[700:760] + "GG" + base[760:840], base[2000:2104][::-1], "AT" * 75, base[1200:1233],
         "".join(rnd.choice("ACGT") for _ in range(150)), base[40:290], "ACG" * 40, base[5:12]]
eng = swb.Engine(0)
n = check_pairs(eng, refs, reads, (5, -3, -4), max_cells=400)
n += check_pairs(eng, refs[:3], reads[:5], (5, -3, -4))
# long reads (int32 wide path, 1-3 bands) and a scores-only / not-fetched / best-hit round trip
long_reads = [base[50:50 + m] for m in (300, 700, 1100)] + ["".join(rnd.choice("ACGT") for _ in range(400))]
n += check_pairs(eng, refs, long_reads, (5, -3, -4), max_cells=200)
n += check_pairs(eng, refs[:4], long_reads[:2] + reads[:2], (2, -1, -2), max_cells=200)
import numpy as np
rs = eng.load_refset(refs)
full = rs.align(reads)
so = rs.align(reads, scores_only=True)
late = rs.align(reads, fetch=False); late.fetch()
assert (so.scores == full.scores).all() and (late.scores == full.scores).all()
assert (so.best_hits[:, :2] == full.best_hits[:, :2]).all() and (late.best_hits == full.best_hits).all()
assert (late.cell_offsets == full.cell_offsets).all() and (late.cells == full.cells).all() and (late.op_lens == full.op_lens).all()
best = full.best_hits
for q in range(len(reads)):
    col = full.scores[:, q]
    assert best[q, 0] == col.max() and best[q, 1] == int(np.argmax(col))
full.free(); so.free(); late.free(); rs.free()
eng.close()
print("checked", n)
""" % ROOT


@pytest.mark.parametrize("switch", ["SWB_NO_BIAS_FILL", "SWB_NO_SUBSAMPLE", "SWB_NO_TILE_TRACE",
                                    "SWB_NO_BIAS_FILL,SWB_NO_TILE_TRACE",
                                    # the reference set cut into 3 / 6 parts (part k's results cross PCIe while part k + 1
                                    # computes): the stitched ABI arrays must be the single-part ones
                                    "SWB_REF_PARTS=3", "SWB_REF_PARTS=6,SWB_NO_TILE_TRACE",
                                    # every wide-path geometry and traceback grouping on the same inputs
                                    "SWB_WIDE_KL=8", "SWB_WIDE_KL=32,SWB_WIDE_TRACE_G=1", "SWB_WIDE_KL=16,SWB_WIDE_TRACE_G=32",
                                    "SWB_WIDE_KL=32,SWB_WIDE_TRACE_G=128", "SWB_WIDE_KL=16,SWB_WIDE_TRACE_G=128,SWB_WIDE_PIPE=2",
                                    "SWB_WIDE_KL=8,SWB_WIDE_TRACE_G=128,SWB_WIDE_PIPE=4", "SWB_WIDE_NO_PROFILE"])
def test_forced_fallback_kernels_are_exact(switch):
    env = dict(os.environ)
    for s in switch.split(","):
        name, _, val = s.partition("=")
        env[name] = val or "1"
    p = subprocess.run([sys.executable, "-c", SCRIPT], env=env, cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-4000:]
    assert "checked" in p.stdout
