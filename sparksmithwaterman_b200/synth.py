"""Deterministic synthetic data shaped like the reference's RefSeq release
(README.md:36-40 of the reference: mean 2,160 bp, median 1,609 bp per sequence).

Lengths are lognormal: mu = ln 1609, sigma = sqrt(2 ln(2160/1609)), clipped to
[50, 200000]; bases iid uniform ACGT; PRNG numpy PCG64 with the seeds below, so the
oracle and the GPU path always see identical bytes."""
from __future__ import annotations

import math
from typing import List, Tuple

import numpy as np

REF_SEED = 20151001
READ_SEED = 20151002
MU = math.log(1609.0)
SIGMA = math.sqrt(2.0 * math.log(2160.0 / 1609.0))
_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def ref_lengths(n_refs: int, seed: int = REF_SEED, lo: int = 50, hi: int = 200_000) -> np.ndarray:
    rng = np.random.Generator(np.random.PCG64(seed))
    L = np.rint(np.exp(rng.normal(MU, SIGMA, size=n_refs))).astype(np.int64)
    return np.clip(L, lo, hi)


def make_refs(n_refs: int, seed: int = REF_SEED, lo: int = 50, hi: int = 200_000) -> List[bytes]:
    L = ref_lengths(n_refs, seed, lo, hi)
    rng = np.random.Generator(np.random.PCG64(seed + 1))
    flat = _ACGT[rng.integers(0, 4, size=int(L.sum()), dtype=np.uint8)]
    out, p = [], 0
    for n in L:
        out.append(flat[p:p + int(n)].tobytes())
        p += int(n)
    return out


def make_reads(n_reads: int, read_len: int, refs: List[bytes], seed: int = READ_SEED,
               planted_frac: float = 0.5, sub_rate: float = 0.02, indel_rate: float = 0.005) -> List[bytes]:
    """Half the reads are substrings of a random reference with substitutions and
    single-base indels, half are iid random."""
    rng = np.random.Generator(np.random.PCG64(seed))
    out = []
    for _ in range(n_reads):
        if refs and rng.random() < planted_frac:
            r = refs[int(rng.integers(0, len(refs)))]
            if len(r) > read_len + 8:
                s = int(rng.integers(0, len(r) - read_len - 8))
                src = np.frombuffer(r[s:s + read_len + 8], dtype=np.uint8)
                buf = []
                k = 0
                while len(buf) < read_len and k < len(src):
                    u = rng.random()
                    if u < indel_rate / 2:
                        k += 1                                   # deletion from the read
                        continue
                    if u < indel_rate:
                        buf.append(int(_ACGT[int(rng.integers(0, 4))]))   # insertion into the read
                        continue
                    b = int(src[k]); k += 1
                    if rng.random() < sub_rate:
                        b = int(_ACGT[int(rng.integers(0, 4))])
                    buf.append(b)
                while len(buf) < read_len:
                    buf.append(int(_ACGT[int(rng.integers(0, 4))]))
                out.append(bytes(buf[:read_len]))
                continue
        out.append(_ACGT[rng.integers(0, 4, size=read_len, dtype=np.uint8)].tobytes())
    return out


def workload(n_reads: int, read_len: int, n_refs: int, ref_seed: int = REF_SEED,
             read_seed: int = READ_SEED) -> Tuple[List[bytes], List[bytes]]:
    refs = make_refs(n_refs, ref_seed)
    return refs, make_reads(n_reads, read_len, refs, read_seed)
