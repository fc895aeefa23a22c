"""Bloom filter size estimation for `Database.set_configs` (reference: core.py:470-533, which shells out to ntCard).

The reference sets `filter_size = F0` (ntCard's ESTIMATE of the number of distinct canonical k-mers over all input
genomes; its histogram parse never subtracts the low-abundance classes), plus `filter_expand_by`
percent.  Here F0 comes from the hot path itself: every genome is hashed by K1 into ONE shared Bloom filter of m bits
with `modulus = S * m` -- HowDeSBT's own rule "set the bit only if h % modulus < num_bits" keeps a uniform 1/S sample
of the hash space -- and the number of distinct k-mers follows from the set bits X by linear counting,
`F0 ~= S * -m * ln(1 - X/m)`.  With the load kept below 1/8 the estimator's relative standard error is far below
ntCard's own (a few per cent); it is an estimate all the same, not a bit-exact stand-in for ntCard's number.
"""
from __future__ import annotations

import math
import os
from typing import Callable, Sequence, Tuple

EST_BITS = 1 << 30          # 128 MiB counting filter
EST_MAX_LOAD = 1.0 / 8.0    # distinct k-mers per bit the sampling factor aims below


def linear_count(set_bits: int, num_bits: int) -> float:
    """Expected number of distinct items that set `set_bits` of `num_bits` bits with one hash function."""
    if set_bits < 0 or set_bits > num_bits or num_bits <= 0:
        raise ValueError("set_bits must be within [0, num_bits]")
    if set_bits == num_bits:
        raise ValueError("the counting filter is saturated")
    return -float(num_bits) * math.log1p(-float(set_bits) / float(num_bits))


def sampling_factor(total_bases: int, num_bits: int = EST_BITS) -> int:
    """S such that at most total_bases / S distinct k-mers land in the counting filter at <= EST_MAX_LOAD per bit."""
    return max(1, int(math.ceil(total_bases / (EST_MAX_LOAD * num_bits))))


def estimate_distinct_kmers(filepaths: Sequence[str], kmer_size: int,
                            count_union_bits: Callable[[Sequence[str], int, int, int], int],
                            num_bits: int = EST_BITS) -> Tuple[int, int]:
    """(F0 estimate, sampling factor).  `count_union_bits(filepaths, k, num_bits, modulus)` returns the popcount of the
    union filter of all files (backend.count_union_bits on the GPU)."""
    total = sum(os.path.getsize(p) for p in filepaths)  # upper bound on the number of bases
    S = sampling_factor(total, num_bits)
    X = int(count_union_bits(filepaths, kmer_size, num_bits, S * num_bits))
    return int(round(S * linear_count(X, num_bits))), S


def filter_size_from_f0(f0: int, filter_expand_by: float = None) -> int:
    """core.py:521-529: filter_size = F0, expanded by int(filter_size * pct / 100)."""
    if f0 <= 0:
        raise ValueError("The input genomes contain no k-mer of the requested size")
    size = int(f0)
    if filter_expand_by:
        size += int((size * filter_expand_by) / 100.0)
    return size
