"""ctypes front-end and pure-Python twin of oracle/msbt_oracle.c.

TEST INFRASTRUCTURE ONLY -- PARITY UNPINNED (see the header of msbt_oracle.c).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module.  The shipped package (metasbt_b200/) never does.

Two independent CPU restatements live here so that they can check each other:

* ``lib()``      -- the C restatement compiled with gcc (fast enough for Mbp inputs);
* ``py_*``       -- pure-Python integer arithmetic for tiny cases only.

Reference call sites restated (all in /root/reference/metasbt/core.py):
makebf :3920-3931, bfoperate --or :3851-3857, bfdistance :1747-1774, dumpbf :3588-3605,
query :2039-2089, ANI formula :1816-1829.
"""

from __future__ import annotations

import ctypes
import math
import os
import subprocess
from typing import Iterable, List, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "msbt_oracle.c")
_SO = os.path.join(_HERE, "_build", "libmsbt_oracle.so")
_LIB = None

MASK64 = (1 << 64) - 1


def build(force: bool = False) -> str:
    """Compile the C restatement with gcc into oracle/_build/ (git-ignored)."""
    if force or not os.path.isfile(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        subprocess.check_call(
            ["gcc", "-O2", "-Wall", "-Wextra", "-shared", "-fPIC", "-o", _SO, _SRC]
        )
    return _SO


def lib() -> ctypes.CDLL:
    global _LIB
    if _LIB is None:
        if not os.path.isfile(_SO) or (
            os.path.isfile(_SRC) and os.path.getmtime(_SO) < os.path.getmtime(_SRC)
        ):
            build()
        L = ctypes.CDLL(_SO)
        u64, u32, i64 = ctypes.c_uint64, ctypes.c_uint32, ctypes.c_int64
        vp, cp, sz = ctypes.c_void_p, ctypes.c_char_p, ctypes.c_size_t
        L.orc_murmur3_x64_128.argtypes = [cp, ctypes.c_int, u32, vp]
        L.orc_murmur3_x64_128.restype = None
        L.orc_kmer_hash.argtypes = [cp, ctypes.c_int, u64, ctypes.POINTER(ctypes.c_int)]
        L.orc_kmer_hash.restype = u64
        L.orc_makebf.argtypes = [cp, sz, ctypes.c_int, u64, u64, u64, u32, vp]
        L.orc_makebf.restype = i64
        L.orc_makebf_rolling.argtypes = [cp, sz, ctypes.c_int, u64, u64, u64, vp]
        L.orc_makebf_rolling.restype = i64
        L.orc_positions_distinct.argtypes = [cp, sz, ctypes.c_int, u64, u64, u64, ctypes.POINTER(vp)]
        L.orc_positions_distinct.restype = i64
        L.orc_free.argtypes = [vp]
        L.orc_free.restype = None
        L.orc_query_hits.argtypes = [vp, u64, vp, u64, vp]
        L.orc_query_hits.restype = None
        L.orc_bf_or.argtypes = [vp, vp, u64, u64]
        L.orc_bf_or.restype = None
        for name in ("orc_bf_and_popcount", "orc_bf_or_popcount"):
            getattr(L, name).argtypes = [vp, vp, u64]
            getattr(L, name).restype = u64
        L.orc_bf_popcount.argtypes = [vp, u64]
        L.orc_bf_popcount.restype = u64
        _LIB = L
    return _LIB


def _words(num_bits: int) -> int:
    return (num_bits + 63) // 64


def _ptr(a: np.ndarray) -> int:
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data


# ------------------------------------------------------------------ C restatement


def murmur3_x64_128(data: bytes, seed: int = 0) -> Tuple[int, int]:
    out = np.zeros(2, dtype=np.uint64)
    lib().orc_murmur3_x64_128(data, len(data), seed & 0xFFFFFFFF, _ptr(out))
    return int(out[0]), int(out[1])


def kmer_hash(kmer: str, seed: int = 0) -> int:
    ok = ctypes.c_int(0)
    h = lib().orc_kmer_hash(kmer.encode(), len(kmer), seed, ctypes.byref(ok))
    if not ok.value:
        raise ValueError(f"not an ACGT k-mer of length 1..64: {kmer!r}")
    return int(h)


def set_spec(key_bytes: str = "quarter", digest_half: int = 0, canonical: str = "min") -> None:
    """The oracle's mirror of msbt_hash_spec (include/msbt.h): switches for the HowDeSBT-side choices that are not pinned
    against a real binary.  Process-wide; call set_spec() with no arguments to go back to the default (App. A)."""
    L = lib()
    L.orc_set_spec.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int]
    L.orc_set_spec.restype = None
    L.orc_set_spec({"quarter": 0, "words": 1}[key_bytes], int(digest_half), {"min": 0, "forward": 1}[canonical])


def makebf(fasta: bytes, k: int, num_bits: int, min_abund: int = 1, seed: int = 0,
           modulus: int | None = None, rolling: bool = False) -> np.ndarray:
    """howdesbt makebf --k --min --bits --hashes=1 --seed=seed,0 -> u64 words of the filter."""
    modulus = num_bits if modulus is None else modulus
    words = np.zeros(_words(num_bits), dtype=np.uint64)
    if rolling:
        if min_abund > 1:
            raise ValueError("rolling variant is min_abund=1 only")
        rc = lib().orc_makebf_rolling(fasta, len(fasta), k, seed, modulus, num_bits, _ptr(words))
    else:
        rc = lib().orc_makebf(fasta, len(fasta), k, seed, modulus, num_bits, min_abund, _ptr(words))
    if rc < 0:
        raise ValueError("invalid makebf arguments")
    return words


def positions_distinct(fasta: bytes, k: int, num_bits: int, seed: int = 0,
                       modulus: int | None = None) -> np.ndarray:
    """Sorted distinct hash positions of all k-mers of the file (query --distinctkmers)."""
    modulus = num_bits if modulus is None else modulus
    p = ctypes.c_void_p()
    n = lib().orc_positions_distinct(fasta, len(fasta), k, seed, modulus, num_bits, ctypes.byref(p))
    if n < 0:
        raise ValueError("invalid arguments")
    try:
        arr = np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_uint64)), shape=(max(n, 1),))
        return arr[:n].copy()
    finally:
        lib().orc_free(p)


def query_hits(positions: np.ndarray, filters: Sequence[np.ndarray]) -> np.ndarray:
    positions = np.ascontiguousarray(positions, dtype=np.uint64)
    ptrs = (ctypes.c_void_p * len(filters))(*[_ptr(f) for f in filters])
    hits = np.zeros(len(filters), dtype=np.uint64)
    lib().orc_query_hits(_ptr(positions) if len(positions) else None, len(positions),
                         ctypes.cast(ptrs, ctypes.c_void_p), len(filters), _ptr(hits))
    return hits


def bf_or(filters: Sequence[np.ndarray]) -> np.ndarray:
    out = np.zeros_like(filters[0])
    ptrs = (ctypes.c_void_p * len(filters))(*[_ptr(f) for f in filters])
    lib().orc_bf_or(_ptr(out), ctypes.cast(ptrs, ctypes.c_void_p), len(filters), len(out))
    return out


def bf_popcount(a: np.ndarray) -> int:
    return int(lib().orc_bf_popcount(_ptr(a), len(a)))


def bf_and_popcount(a: np.ndarray, b: np.ndarray) -> int:
    return int(lib().orc_bf_and_popcount(_ptr(a), _ptr(b), len(a)))


def bf_or_popcount(a: np.ndarray, b: np.ndarray) -> int:
    return int(lib().orc_bf_or_popcount(_ptr(a), _ptr(b), len(a)))


# ------------------------------------------------------------------ host formulas (core.py)


def ani_distance(intersect: int, union: int, kmer_size: int) -> float:
    """core.py:1816-1829 (raises ZeroDivisionError when union == 0, like the reference)."""
    jaccard_index = round(intersect / union, 5)
    if jaccard_index == 0.0:
        return 1.0
    d = 1 - (1 + (1 / kmer_size) * math.log((2 * jaccard_index) / (1 + jaccard_index)))
    return 1.0 if d > 1.0 else d


def cluster_topology(filters: Sequence[np.ndarray]) -> List[Tuple[int, int]]:
    """howdesbt cluster --bits=<all> --keepallnodes (core.py:3800-3808), brute force: while more than one
    node is active, recompute ALL pairwise Hamming distances between active nodes, merge the closest pair (ties: smaller
    node ids; leaves are 0..n-1, new nodes follow in creation order) into their OR.  -> [(depth, node id)] in pre-order,
    first child = the lower id of a merged pair.  Tie-breaking and numbering are UNPINNED (no howdesbt binary)."""
    nodes = {i: np.asarray(f, dtype=np.uint64) for i, f in enumerate(filters)}
    n = len(nodes)
    if n == 0:
        return []
    children = {}
    nxt = n
    while len(nodes) > 1:
        ids = sorted(nodes)
        best = None
        for a in range(len(ids)):
            for b in range(a + 1, len(ids)):
                d = bf_popcount(np.bitwise_xor(nodes[ids[a]], nodes[ids[b]]))
                if best is None or (d, ids[a], ids[b]) < best:
                    best = (d, ids[a], ids[b])
        _, u, v = best
        nodes[nxt] = np.bitwise_or(nodes.pop(u), nodes.pop(v))
        children[nxt] = (u, v)
        nxt += 1
    out: List[Tuple[int, int]] = []

    def walk(node: int, depth: int) -> None:
        out.append((depth, node))
        if node in children:
            walk(children[node][0], depth + 1)
            walk(children[node][1], depth + 1)
    walk(next(iter(nodes)), 0)
    return out


# ------------------------------------------------------------------ pure-Python twin (tiny cases)


def _rotl(x: int, r: int) -> int:
    return ((x << r) | (x >> (64 - r))) & MASK64


def _fmix(k: int) -> int:
    k ^= k >> 33
    k = (k * 0xFF51AFD7ED558CCD) & MASK64
    k ^= k >> 33
    k = (k * 0xC4CEB9FE1A85EC53) & MASK64
    k ^= k >> 33
    return k


def py_murmur3_x64_128(data: bytes, seed: int = 0) -> Tuple[int, int]:
    c1, c2 = 0x87C37B91114253D5, 0x4CF5AD432745937F
    h1 = h2 = seed & 0xFFFFFFFF
    n = len(data)
    nb = n // 16
    for i in range(nb):
        k1 = int.from_bytes(data[16 * i:16 * i + 8], "little")
        k2 = int.from_bytes(data[16 * i + 8:16 * i + 16], "little")
        k1 = (k1 * c1) & MASK64; k1 = _rotl(k1, 31); k1 = (k1 * c2) & MASK64; h1 ^= k1
        h1 = _rotl(h1, 27); h1 = (h1 + h2) & MASK64; h1 = (h1 * 5 + 0x52DCE729) & MASK64
        k2 = (k2 * c2) & MASK64; k2 = _rotl(k2, 33); k2 = (k2 * c1) & MASK64; h2 ^= k2
        h2 = _rotl(h2, 31); h2 = (h2 + h1) & MASK64; h2 = (h2 * 5 + 0x38495AB5) & MASK64
    tail = data[16 * nb:]
    if len(tail) > 8:
        k2 = int.from_bytes(tail[8:], "little")
        k2 = (k2 * c2) & MASK64; k2 = _rotl(k2, 33); k2 = (k2 * c1) & MASK64; h2 ^= k2
    if len(tail) > 0:
        k1 = int.from_bytes(tail[:8], "little")
        k1 = (k1 * c1) & MASK64; k1 = _rotl(k1, 31); k1 = (k1 * c2) & MASK64; h1 ^= k1
    h1 ^= n; h2 ^= n
    h1 = (h1 + h2) & MASK64; h2 = (h2 + h1) & MASK64
    h1 = _fmix(h1); h2 = _fmix(h2)
    h1 = (h1 + h2) & MASK64; h2 = (h2 + h1) & MASK64
    return h1, h2


_CODE = {"A": 0, "C": 1, "G": 2, "T": 3}


def py_canonical(kmer: str) -> int:
    """Smaller of the k-mer and its reverse complement as base-4 integers, first base most significant."""
    codes = [_CODE[c] for c in kmer.upper()]
    f = 0
    for c in codes:
        f = (f << 2) | c
    r = 0
    for c in reversed(codes):
        r = (r << 2) | (3 - c)
    return min(f, r)


def py_kmer_hash(kmer: str, seed: int = 0, key_bytes: str = "quarter", digest_half: int = 0, canonical: str = "min") -> int:
    """Independent pure-Python statement of the k-mer hash, with the spec switches of set_spec() as arguments."""
    k = len(kmer)
    if canonical == "min":
        value = py_canonical(kmer)
    else:
        value = 0
        for c in kmer.upper():
            value = (value << 2) | _CODE[c]
    length = (k + 3) // 4 if key_bytes == "quarter" else (8 if k <= 32 else 16)
    return py_murmur3_x64_128(value.to_bytes(16, "little")[:length], seed)[digest_half]


def py_fasta_records(fasta: bytes) -> List[str]:
    recs: List[List[str]] = []
    for line in fasta.decode("latin-1").split("\n"):
        if line.startswith(">"):
            recs.append([])
        else:
            if not recs:
                recs.append([])
            recs[-1].append("".join(line.split()))
    return ["".join(r) for r in recs]


def py_kmers(fasta: bytes, k: int) -> Iterable[str]:
    for rec in py_fasta_records(fasta):
        for i in range(len(rec) - k + 1):
            w = rec[i:i + k].upper()
            if all(c in _CODE for c in w):
                yield w


def py_makebf(fasta: bytes, k: int, num_bits: int, min_abund: int = 1, seed: int = 0,
              modulus: int | None = None) -> np.ndarray:
    modulus = num_bits if modulus is None else modulus
    counts: dict = {}
    for w in py_kmers(fasta, k):
        c = py_canonical(w)
        counts[c] = counts.get(c, 0) + 1
    bits = 0
    for c, n in counts.items():
        if n >= max(min_abund, 1):
            pos = py_murmur3_x64_128(c.to_bytes(16, "little")[: (k + 3) // 4], seed)[0] % modulus
            if pos < num_bits:
                bits |= 1 << pos
    nw = _words(num_bits)
    return np.frombuffer(bits.to_bytes(nw * 8, "little"), dtype=np.uint64).copy()


def py_positions_distinct(fasta: bytes, k: int, num_bits: int, seed: int = 0,
                          modulus: int | None = None) -> np.ndarray:
    modulus = num_bits if modulus is None else modulus
    s = set()
    for w in py_kmers(fasta, k):
        pos = py_kmer_hash(w, seed) % modulus
        if pos < num_bits:
            s.add(pos)
    return np.array(sorted(s), dtype=np.uint64)


# ---- k-mer size search (metasbt_b200/kopt.py): the three-step method of Zhang et al. 2017 that `kitsune kopt` implements
# (reference call site: core.py:439-468; kitsune itself is absent and unpinned).  Exact, dictionary- and set-based.

def py_canonical_str(w: str) -> str:
    r = w[::-1].translate(str.maketrans("ACGT", "TGCA"))
    return min(w, r)


def py_word_counts(fasta: bytes, l: int) -> dict:
    counts: dict = {}
    for w in py_kmers(fasta, l):
        c = py_canonical_str(w)
        counts[c] = counts.get(c, 0) + 1
    return counts


def py_relative_entropy(fasta: bytes, l: int) -> float:
    """RE(l) = sum_w f_l(w) log2( f_l(w) f_{l-2}(w[1:-1]) / (f_{l-1}(w[:-1]) f_{l-1}(w[1:])) ) over canonical words."""
    import math
    cl, c1, c2 = py_word_counts(fasta, l), py_word_counts(fasta, l - 1), py_word_counts(fasta, l - 2)
    nl, n1, n2 = sum(cl.values()), sum(c1.values()), sum(c2.values())
    if nl == 0:
        return 0.0
    total = 0.0
    for w, c in cl.items():
        f = c / nl
        fp = c1[py_canonical_str(w[:-1])] / n1
        fs = c1[py_canonical_str(w[1:])] / n1
        fm = c2[py_canonical_str(w[1:-1])] / n2
        total += f * math.log2(f * fm / (fp * fs))
    return total


def py_kopt(fastas: Sequence[bytes], k_max: int, k_min: int = 4, cre_cutoff: float = 0.1, acf_cutoff: float = 0.1) -> dict:
    """{'k', 'k_lo', 'k_hi', 'acf', 'ofc', 'cre_limits'} with exact set sizes."""
    k_min = max(3, min(k_min, k_max))
    limits = []
    for fa in fastas:
        re = {l: py_relative_entropy(fa, l) for l in range(k_min, k_max + 1)}
        cre, acc = {}, 0.0
        for k in range(k_max, k_min - 1, -1):
            acc += re[k]
            cre[k] = acc
        top = max(cre.values())
        lim = k_max
        if top <= 0.0:
            lim = k_min
        else:
            for k in range(k_min, k_max + 1):
                if cre[k] <= cre_cutoff * top:
                    lim = k
                    break
        limits.append(lim)
    k_lo = max(limits)
    out = {"cre_limits": limits, "k_lo": k_lo}
    n = len(fastas)
    if n < 2:
        out["k"] = k_lo
        return out
    acf, ofc = {}, {}
    for k in range(k_min, k_max + 1):
        sets = [set(py_word_counts(fa, k)) for fa in fastas]
        acf[k] = sum(len(sets[i] & sets[j]) for i in range(n) for j in range(i + 1, n)) / (n * (n - 1) / 2)
        seen, twice = set(), set()
        for s in sets:
            twice |= seen & s
            seen |= s
        ofc[k] = float(len(twice))
    top = max(acf.values())
    k_hi = k_min if top <= 0 else max(k for k in acf if acf[k] >= acf_cutoff * top)
    best = k_lo
    if k_lo <= k_hi:
        for k in range(k_lo, k_hi + 1):
            if ofc[k] > ofc[best]:
                best = k
    out.update({"acf": acf, "ofc": ofc, "k_hi": k_hi, "k": best})
    return out
