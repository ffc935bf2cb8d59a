"""sparksmithwaterman_b200 -- B200-native (sm_100a) drop-in for the `sw` hot path of
elizabethfong/SparkSmithWaterman: linear-gap Smith-Waterman fill + traceback of every
maximum-score cell over all read x reference pairs, behind a plain C ABI
(include/swb200.h).  The CUDA library is mandatory: nothing here computes on the CPU."""
from .engine import AlignResult, DEFAULT_SCORES, Engine, Reads, RefSet  # noqa: F401
from ._ffi import SwbError  # noqa: F401

__all__ = ["Engine", "RefSet", "Reads", "AlignResult", "SwbError", "DEFAULT_SCORES"]
