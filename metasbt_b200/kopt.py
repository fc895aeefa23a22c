"""k-mer size search for `Database.set_configs` (reference: core.py:430-468, which shells out to
`kitsune kopt --filenames <list> --k-max K --canonical --fast --in-memory` and takes the last token of its output file).

kitsune (github.com/natapol/kitsune) is third-party, not under the reference tree and not version-pinned by it
(`requirements.txt` / `setup.py` do not list it; core.py:39 only names the executable).  This module restates the
published method the tool implements -- the three-step approach of Zhang, Jun, Leuze, Ussery & Nookaew, "Viral
phylogenomics using an alignment-free method: a three-step approach to determine optimal length of k-mer", Sci. Rep. 7,
40712 (2017) -- over canonical k-mers (`--canonical`).  **Parity with the real tool is unpinned**: the cut-offs and the
way the three steps are combined follow the paper and the tool's documented defaults (10 % for CRE and ACF) as recalled,
there is no golden vector to check them against.

  1. CRE, cumulative relative entropy, per genome: RE(l) is the Kullback-Leibler divergence between the observed l-mer
     frequencies f_l and the ones an (l-2)-order Markov model predicts from shorter words,
         f^_l(w) = f_{l-1}(w[:-1]) * f_{l-1}(w[1:]) / f_{l-2}(w[1:-1]),      RE(l) = sum_w f_l(w) log2(f_l(w) / f^_l(w)),
     and CRE(k) = sum_{l >= k} RE(l) (to k_max): the information lost by using words shorter than k.  The lower limit
     of a genome is the smallest k whose CRE is at most 10 % of the largest CRE; over genomes the largest limit counts.
  2. ACF, average number of common features: the mean over genome pairs of the number of distinct k-mers they share.
     The upper limit is the largest k whose ACF is at least 10 % of the largest ACF.
  3. OFC, observed feature occurrence: the number of distinct k-mers seen in two or more genomes; the optimum is the k
     inside [lower, upper] with the most of them (the smallest such k on a tie; the lower limit when the range is empty).

Where the work runs: step 1 is exact word counting of one genome at a time (sort-based `torch.unique` on 64-bit keys
on the backend's device; windows never span records or non-ACGT characters); steps 2 and 3 touch every pair of genomes
and go through the hot-path kernels: K1 builds one Bloom filter per genome and k over a uniform 1/S sample of the hash
space (`modulus = S * num_bits`, the rule estimate.py uses), K3 all-pairs gives the pairwise intersections and the
`msbt_bf_occurrence` pass the bits set in >= 1 and >= 2 genomes; set sizes follow by linear counting.  At most
`max_genomes` genomes (evenly spaced over the sorted input list) take part -- kitsune loads every genome in memory.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

K_MIN = 4                 # kitsune's default first k
K_LIMIT = 32              # 64-bit keys
CRE_CUTOFF = 0.1
ACF_CUTOFF = 0.1
SET_BITS = 1 << 27        # bits per genome filter in steps 2 and 3
SET_MAX_LOAD = 1.0 / 8.0  # distinct k-mers per bit (all genomes together) the sampling factor aims below
MAX_GENOMES = 64

_LUT = np.full(256, 4, dtype=np.uint8)
for _i, _c in enumerate("ACGT"):
    _LUT[ord(_c)] = _i
    _LUT[ord(_c.lower())] = _i


def fasta_codes(text: bytes) -> np.ndarray:
    """FASTA text -> one uint8 code per base (A, C, G, T = 0..3, anything else 4) with a 4 between records; header
    lines and line breaks are dropped (the record handling of oracle.py_fasta_records)."""
    a = np.frombuffer(text, dtype=np.uint8)
    if a.size == 0:
        return np.zeros(0, dtype=np.uint8)
    nl = a == 10
    starts = np.concatenate((np.zeros(1, dtype=np.int64), np.flatnonzero(nl) + 1))
    starts = starts[starts < a.size]
    header_starts = starts[a[starts] == ord(">")]
    line_id = np.cumsum(nl) - nl                      # index of the line every byte belongs to
    is_header_line = np.zeros(int(line_id[-1]) + 1, dtype=bool)
    is_header_line[line_id[header_starts]] = True
    in_header = is_header_line[line_id]
    codes = _LUT[a]
    codes[header_starts] = 4                          # one break per header
    keep = ~in_header & ~nl & (a != 13) & (a != 32) & (a != 9)
    keep[header_starts] = True
    return codes[keep]


def _log2_window_counts(keys: torch.Tensor, valid: torch.Tensor) -> Tuple[torch.Tensor, int]:
    """log2 of the number of valid windows that carry the same key as window i (0 for invalid windows), and the number
    of valid windows."""
    out = torch.zeros(keys.numel(), dtype=torch.float64, device=keys.device)
    idx = torch.nonzero(valid).squeeze(1)
    if idx.numel() == 0:
        return out, 0
    _, inverse, counts = torch.unique(keys[idx], return_inverse=True, return_counts=True)
    out[idx] = torch.log2(counts.to(torch.float64))[inverse]
    return out, int(idx.numel())


def relative_entropies(codes: np.ndarray, k_max: int, device: torch.device | str) -> Dict[int, float]:
    """RE(l) for l = 3..k_max of one genome given as fasta_codes(); canonical words, exact counts."""
    if k_max > K_LIMIT:
        raise ValueError(f"word lengths above {K_LIMIT} are not supported")
    c = torch.from_numpy(np.ascontiguousarray(codes)).to(device).to(torch.int64)
    G = c.numel()
    bad = (c > 3).to(torch.int64)
    cum = torch.zeros(G + 1, dtype=torch.int64, device=c.device)
    torch.cumsum(bad, 0, out=cum[1:])
    c = c.clamp(max=3)
    comp = 3 - c
    sign = torch.tensor(-(1 << 63), dtype=torch.int64, device=c.device)
    fwd = torch.zeros(G, dtype=torch.int64, device=c.device)   # fwd[i], rc[i]: the window of the current length at i
    rc = torch.zeros(G, dtype=torch.int64, device=c.device)
    logs: Dict[int, torch.Tensor] = {}
    totals: Dict[int, int] = {}
    res: Dict[int, float] = {}
    for l in range(1, k_max + 1):
        n = G - l + 1
        if n <= 0:
            totals[l] = 0
            logs[l] = torch.zeros(0, dtype=torch.float64, device=c.device)
            if l >= 3:
                res[l] = 0.0
            continue
        fwd = (fwd[:n] << 2) | c[l - 1: l - 1 + n]
        rc = rc[:n] | (comp[l - 1: l - 1 + n] << (2 * (l - 1)))
        canon = torch.where((fwd ^ sign) < (rc ^ sign), fwd, rc) if l == 32 else torch.minimum(fwd, rc)
        valid = (cum[l: l + n] - cum[:n]) == 0
        logs[l], totals[l] = _log2_window_counts(canon, valid)
        if l >= 3:
            if totals[l] == 0:
                res[l] = 0.0
            else:
                Nl, N1, N2 = totals[l], totals[l - 1], totals[l - 2]
                # per valid window i: log2( f_l(w) f_{l-2}(mid) / (f_{l-1}(pre) f_{l-1}(suf)) )
                term = logs[l] + logs[l - 2][1: n + 1] - logs[l - 1][:n] - logs[l - 1][1: n + 1]
                s = float(torch.where(valid, term, torch.zeros_like(term)).sum())
                res[l] = s / Nl + (2.0 * math.log2(N1) - math.log2(Nl) - math.log2(N2))
            logs.pop(l - 2, None)
    return res


def cre_from_re(re: Dict[int, float], k_min: int, k_max: int) -> Dict[int, float]:
    out, acc = {}, 0.0
    for k in range(k_max, k_min - 1, -1):
        acc += re.get(k, 0.0)
        out[k] = acc
    return out


def cre_lower_limit(cre: Dict[int, float], cutoff: float = CRE_CUTOFF) -> int:
    ks = sorted(cre)
    top = max(cre.values())
    if top <= 0.0:
        return ks[0]
    for k in ks:
        if cre[k] <= cutoff * top:
            return k
    return ks[-1]


def acf_upper_limit(acf: Dict[int, float], cutoff: float = ACF_CUTOFF) -> int:
    ks = sorted(acf)
    top = max(acf.values())
    if top <= 0.0:
        return ks[0]
    return max(k for k in ks if acf[k] >= cutoff * top)


def choose(k_lo: int, k_hi: int, ofc: Dict[int, float]) -> int:
    if k_lo > k_hi:
        return k_lo
    best = k_lo
    for k in range(k_lo, k_hi + 1):
        if ofc[k] > ofc[best]:
            best = k
    return best


# ---- set sizes from Bloom filter popcounts (steps 2 and 3)

def _linear(bits: float, num_bits: int) -> float:
    if bits >= num_bits:
        raise ValueError("a counting filter is saturated: raise num_bits")
    return -float(num_bits) * math.log1p(-float(bits) / float(num_bits))


def common_features(pop_i: int, pop_j: int, inter: int, num_bits: int, sample: int) -> float:
    """|K_i and K_j| from the popcounts of the two filters and of their AND (the OR has pop_i + pop_j - inter bits)."""
    est = _linear(pop_i, num_bits) + _linear(pop_j, num_bits) - _linear(pop_i + pop_j - inter, num_bits)
    return max(0.0, sample * est)


def shared_features(ge1: int, ge2: int, num_bits: int, sample: int) -> float:
    """Distinct k-mers present in >= 2 genomes from the number of bits set in >= 1 (`ge1`) and >= 2 (`ge2`) genome
    filters.  With Poisson bit occupancy, T = -ln(1 - p1) is the density of all distinct k-mers and a bit counts as
    'twice' either through a shared k-mer or through k-mers of two different genomes colliding; to first order
    (1 - p2) = e^-T (1 + a) with a the density of single-genome k-mers, hence shared = T - a."""
    p1, p2 = ge1 / float(num_bits), ge2 / float(num_bits)
    if p1 >= 1.0:
        raise ValueError("a counting filter is saturated: raise num_bits")
    T = -math.log1p(-p1)
    a = (p1 - p2) / (1.0 - p1)
    return max(0.0, sample * num_bits * (T - a))


def pick_genomes(filepaths: Sequence[str], max_genomes: int = MAX_GENOMES) -> List[str]:
    paths = sorted(filepaths)
    if len(paths) <= max_genomes:
        return paths
    step = len(paths) / float(max_genomes)
    return [paths[int(i * step)] for i in range(max_genomes)]


def optimal_kmer_size(filepaths: Sequence[str], k_max: int, backend, k_min: int = K_MIN, num_bits: int = SET_BITS,
                      max_genomes: int = MAX_GENOMES, cre_cutoff: float = CRE_CUTOFF, acf_cutoff: float = ACF_CUTOFF,
                      details: Optional[dict] = None) -> int:
    """The k `kitsune kopt --k-max k_max --canonical` stands for (see the module docstring).  `backend` provides
    `device` (where the word counting runs), `read_fasta(path) -> bytes` and
    `kmer_set_profile(paths, k, num_bits, modulus) -> (pop[n], inter[n][n], ge1, ge2)`."""
    if not filepaths:
        raise ValueError("One or more genome files must be provided in input!")
    if k_max > K_LIMIT:
        raise ValueError(f"The k-mer size search covers sizes up to {K_LIMIT}")
    k_min = max(3, min(k_min, k_max))
    paths = pick_genomes(filepaths, max_genomes)
    # step 1
    lows, total_bases = [], 0
    for p in paths:
        codes = fasta_codes(backend.read_fasta(p))
        total_bases += int(codes.size)
        cre = cre_from_re(relative_entropies(codes, k_max, backend.device), k_min, k_max)
        lows.append(cre_lower_limit(cre, cre_cutoff))
    k_lo = max(lows)
    if details is not None:
        details.update({"genomes": paths, "cre_limits": lows, "k_lo": k_lo})
    if len(paths) < 2:
        return k_lo
    # steps 2 and 3
    sample = max(1, int(math.ceil(total_bases / (SET_MAX_LOAD * num_bits))))
    acf: Dict[int, float] = {}
    ofc: Dict[int, float] = {}
    n = len(paths)
    for k in range(k_min, k_max + 1):
        pop, inter, ge1, ge2 = backend.kmer_set_profile(paths, k, num_bits, sample * num_bits)
        tot = 0.0
        for i in range(n):
            for j in range(i + 1, n):
                tot += common_features(int(pop[i]), int(pop[j]), int(inter[i][j]), num_bits, sample)
        acf[k] = tot / (n * (n - 1) / 2)
        ofc[k] = shared_features(int(ge1), int(ge2), num_bits, sample)
    done = getattr(backend, "kmer_set_profile_done", None)
    if done is not None:
        done()
    k_hi = acf_upper_limit(acf, acf_cutoff)
    best = choose(k_lo, k_hi, ofc)
    if details is not None:
        details.update({"sample": sample, "acf": acf, "ofc": ofc, "k_hi": k_hi, "k": best})
    return best
