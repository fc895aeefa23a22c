"""TEST INFRASTRUCTURE: a checker backend with the interface of metasbt_b200.backend.CudaBackend,
served by the oracle.  It lets the CPU suite exercise the host logic of metasbt_b200.core (file
layout, report, profile descent) without a GPU, and lets GPU tests compare whole databases built by
the two backends.  The shipped package never imports this module."""
import os
from typing import List, Sequence, Tuple

import numpy as np

from metasbt_b200 import bffile
from oracle import oracle as O


class OracleBackend:
    def __init__(self, kmer_size: int, num_bits: int):
        self.k, self.num_bits = int(kmer_size), int(num_bits)
        self._cache = {}

    def _load(self, path: str) -> np.ndarray:
        path = os.path.abspath(path)
        if path not in self._cache:
            header, words = bffile.read_bf(path)
            assert header.num_bits == self.num_bits and header.kmer_size == self.k
            self._cache[path] = words
        return self._cache[path]

    def _store(self, path: str, words: np.ndarray) -> None:
        bffile.write_bf(path, bffile.BloomHeader(kmer_size=self.k, num_bits=self.num_bits), words)
        self._cache[os.path.abspath(path)] = words

    def sketch_many(self, fasta_paths: Sequence[str], out_paths: Sequence[str], min_abund: int) -> None:
        for fa, out in zip(fasta_paths, out_paths):
            self._store(out, O.makebf(open(fa, "rb").read(), self.k, self.num_bits, min_abund))

    def count_union_bits(self, fasta_paths: Sequence[str], kmer_size: int, num_bits: int, modulus: int) -> int:
        acc = None
        for fa in fasta_paths:
            w = O.makebf(open(fa, "rb").read(), kmer_size, num_bits, 1, modulus=modulus, rolling=kmer_size <= 32)
            acc = w if acc is None else O.bf_or([acc, w])
        return 0 if acc is None else O.bf_popcount(acc)

    device = "cpu"

    def read_fasta(self, path: str) -> bytes:
        return open(path, "rb").read()

    def kmer_set_profile(self, fasta_paths: Sequence[str], kmer_size: int, num_bits: int, modulus: int):
        f = [O.makebf(open(fa, "rb").read(), kmer_size, num_bits, 1, modulus=modulus, rolling=kmer_size <= 32) for fa in fasta_paths]
        n = len(f)
        inter = [[O.bf_and_popcount(f[i], f[j]) for j in range(n)] for i in range(n)]
        once, twice = np.zeros_like(f[0]), np.zeros_like(f[0])
        for w in f:
            twice |= once & w
            once |= w
        return [inter[i][i] for i in range(n)], inter, O.bf_popcount(once), O.bf_popcount(twice)

    def union(self, src_paths: Sequence[str], out_path: str) -> None:
        self._store(out_path, O.bf_or([self._load(p) for p in src_paths]))

    def union_many(self, jobs) -> None:
        for srcs, out in jobs:
            self.copy(srcs[0], out) if len(srcs) == 1 else self.union(srcs, out)

    def copy(self, src_path: str, out_path: str) -> None:
        self._store(out_path, self._load(src_path).copy())

    def popcounts(self, paths: Sequence[str]) -> List[int]:
        return [O.bf_popcount(self._load(p)) for p in paths]

    def dist_one_vs_many(self, focus_path: str, target_paths: Sequence[str]) -> List[Tuple[int, int]]:
        a = self._load(focus_path)
        return [(O.bf_and_popcount(a, self._load(p)), O.bf_or_popcount(a, self._load(p))) for p in target_paths]

    def dist_all_pairs(self, paths: Sequence[str]) -> np.ndarray:
        f = [self._load(p) for p in paths]
        return np.array([[O.bf_and_popcount(a, b) for b in f] for a in f], dtype=np.int64)

    def dist_all_pairs_many(self, groups: Sequence[Sequence[str]]) -> List[np.ndarray]:
        return [self.dist_all_pairs(g) for g in groups]

    def dist_pairs(self, pairs):
        return [(O.bf_and_popcount(self._load(a), self._load(b)), O.bf_or_popcount(self._load(a), self._load(b))) for a, b in pairs]

    # the batch interface of profile_many: a "query batch" is just the list of position arrays here
    def query_positions_many(self, fasta_paths: Sequence[str]):
        return [O.positions_distinct(open(p, "rb").read(), self.k, self.num_bits) for p in fasta_paths]

    def query_positions(self, fasta_path: str):
        return self.query_positions_many([fasta_path])

    def query_totals(self, qb) -> List[int]:
        return [len(p) for p in qb]

    def query_hits_many(self, qb, idxs: Sequence[int], leaf_paths: Sequence[str]) -> np.ndarray:
        leaves = [self._load(p) for p in leaf_paths]
        return np.array([[int(x) for x in O.query_hits(qb[i], leaves)] for i in idxs], dtype=np.int64).reshape(len(idxs), len(leaves))

    def fetch_hits(self, pending):
        return list(pending)

    def query_hits(self, positions, leaf_paths: Sequence[str]):
        if not leaf_paths:
            return [], len(positions[0])
        return [int(x) for x in self.query_hits_many(positions, [0], leaf_paths)[0]], len(positions[0])
