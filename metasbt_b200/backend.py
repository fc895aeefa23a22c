"""File-level operators: the five `howdesbt` subcommands MetaSBT spawns, as in-process batch calls
over HBM-resident filters.  core.py talks to this interface only.

    makebf      -> CudaBackend.sketch_many      (core.py:3920-3931)
    bfoperate   -> CudaBackend.union            (core.py:3851-3857)
    dumpbf      -> CudaBackend.popcounts        (core.py:3588-3593)
    bfdistance  -> CudaBackend.dist_one_vs_many / dist_all_pairs   (core.py:1747-1774, 4036-4081)
    query       -> CudaBackend.query_many       (core.py:2039-2047)

Filters are cached on the device by absolute path (LRU, bounded in bytes), so that the repeated
reads the reference pays per spawn (every `bfdistance`/`query` call re-reads every listed file)
happen once.  `.bf` files stay the persistent store and are byte-compatible (bffile.py).
There is no CPU implementation here: without libmsbt.so and a CUDA device, construction fails.
"""

from __future__ import annotations

import os
from collections import OrderedDict
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import bffile, ops


def read_fasta_bytes(path: str) -> bytes:
    with open(path, "rb") as f:
        return f.read()


def read_files(paths: Sequence[str], threads: Optional[int] = None) -> List[bytes]:
    """Read many FASTA files on a thread pool (file reads release the GIL)."""
    from concurrent.futures import ThreadPoolExecutor
    threads = threads or min(32, os.cpu_count() or 1)
    if len(paths) <= 1 or threads <= 1:
        return [read_fasta_bytes(p) for p in paths]
    with ThreadPoolExecutor(max_workers=threads) as ex:
        return list(ex.map(read_fasta_bytes, paths))


class CudaBackend:
    def __init__(self, kmer_size: int, num_bits: int, device: Optional[int] = None,
                 cache_bytes: int = 96 << 30, sketch_batch_bytes: int = 16 << 30):
        self.k = int(kmer_size)
        self.num_bits = int(num_bits)
        self.ctx = ops.default_context(device)
        self.words = ops.filter_words(self.num_bits)
        self.stride = ops.filter_stride_words(self.num_bits)
        self.cache_bytes = cache_bytes
        self.sketch_batch = max(1, sketch_batch_bytes // (self.stride * 8))
        self._cache: "OrderedDict[str, torch.Tensor]" = OrderedDict()
        self._bytes = 0
        # bit-sliced leaf matrices of flat trees that are queried repeatedly (see query_hits)
        self.sliced_min_leaves = 256
        self.sliced_cache_bytes = 48 << 30
        self._sliced: "OrderedDict[Tuple[str, ...], torch.Tensor]" = OrderedDict()
        self._sliced_bytes = 0
        self._tree_queries: Dict[Tuple[str, ...], int] = dict()
        self.sliced_probes = 0  # queries served by the bit-sliced probe (observability / tests)

    # ------------------------------------------------------------------ device filter cache
    def _put(self, path: str, t: torch.Tensor) -> None:
        path = os.path.abspath(path)
        if path in self._cache:
            self._bytes -= self._cache.pop(path).numel() * 8
            self._drop_sliced(path)
        self._cache[path] = t
        self._bytes += t.numel() * 8
        while self._bytes > self.cache_bytes and len(self._cache) > 1:
            _, old = self._cache.popitem(last=False)
            self._bytes -= old.numel() * 8

    def forget(self, path: str) -> None:
        path = os.path.abspath(path)
        t = self._cache.pop(path, None)
        if t is not None:
            self._bytes -= t.numel() * 8
        self._drop_sliced(path)

    def _drop_sliced(self, path: str) -> None:
        """A filter was rewritten: every cached bit-sliced matrix that contains it is stale."""
        for key in [k for k in self._sliced if path in k]:
            self._sliced_bytes -= self._sliced.pop(key).numel() * 4

    def filter(self, path: str) -> torch.Tensor:
        """Device tensor (int64[stride]) of the filter stored at `path` (cached)."""
        path = os.path.abspath(path)
        t = self._cache.get(path)
        if t is not None:
            self._cache.move_to_end(path)
            return t
        header, words = bffile.read_bf(path)
        if header.num_bits != self.num_bits or header.kmer_size != self.k:
            raise ValueError(f"{path}: filter has k={header.kmer_size}, bits={header.num_bits}; "
                             f"database uses k={self.k}, bits={self.num_bits}")
        host = torch.zeros(self.stride, dtype=torch.int64)
        host.numpy().view(np.uint64)[: self.words] = words
        t = host.to(self.ctx.device)
        self._put(path, t)
        return t

    def _write(self, path: str, t: torch.Tensor) -> None:
        header = bffile.BloomHeader(kmer_size=self.k, num_bits=self.num_bits)
        bffile.write_bf(path, header, t[: self.words].cpu().numpy().view(np.uint64))

    # ------------------------------------------------------------------ makebf
    def sketch_many(self, fasta_paths: Sequence[str], out_paths: Sequence[str], min_abund: int) -> None:
        """howdesbt makebf --k --min --bits --hashes=1 --seed=0,0 <fasta> --out=<bf> for every pair."""
        for lo in range(0, len(fasta_paths), self.sketch_batch):
            fa = fasta_paths[lo: lo + self.sketch_batch]
            # raw FASTA text goes host -> device once; parsing and 2-bit packing happen on the GPU (msbt_fasta_pack_device)
            batch = self.ctx.pack_fasta_device(read_files(fa), self.k)
            filt = self.ctx.build_filters(batch, self.k, self.num_bits, min_abund=min_abund)
            outs = out_paths[lo: lo + self.sketch_batch]
            # device -> pinned host copies are queued back to back; the `.bf` files are written by a small thread pool
            # while the next copies are in flight (file writes release the GIL)
            from concurrent.futures import ThreadPoolExecutor
            header = bffile.BloomHeader(kmer_size=self.k, num_bits=self.num_bits)
            with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as pool:
                pending = []
                for i, out in enumerate(outs):
                    row = filt[i].clone()
                    host = torch.empty(self.words, dtype=torch.int64, pin_memory=True)
                    host.copy_(row[: self.words], non_blocking=True)
                    done = torch.cuda.Event()
                    done.record()

                    def write(out=out, host=host, done=done):
                        done.synchronize()
                        bffile.write_bf(out, header, host.numpy().view(np.uint64))
                    pending.append(pool.submit(write))
                    self._put(out, row)
                for job in pending:
                    job.result()

    def sketch_resident(self, fasta_paths: Sequence[str], min_abund: int = 1) -> torch.Tensor:
        """Filters of the given FASTA files left on the device only (query side of `profile`)."""
        batch = self.ctx.pack_fasta_device(read_files(fasta_paths), self.k)
        return self.ctx.build_filters(batch, self.k, self.num_bits, min_abund=min_abund)

    def count_union_bits(self, fasta_paths: Sequence[str], kmer_size: int, num_bits: int, modulus: int) -> int:
        """Popcount of ONE filter of `num_bits` bits that receives every k-mer of every file whose hash, taken modulo
        `modulus`, is below `num_bits` (estimate.py: distinct k-mer counting for the filter size)."""
        acc = None
        groups, size = [[]], 0
        for path in fasta_paths:  # about 1 GiB of FASTA text per pass
            if groups[-1] and size + os.path.getsize(path) > (1 << 30):
                groups.append([])
                size = 0
            groups[-1].append(path)
            size += os.path.getsize(path)
        for group in groups:
            if not group:
                continue
            batch = self.ctx.pack_fasta_device(read_files(group), kmer_size)
            for i in range(batch.n):
                batch.descs[i].filter_index = 0  # all genomes of the batch into one shared filter
            part = self.ctx.build_filters(batch, kmer_size, num_bits, min_abund=1, modulus=modulus)
            acc = part[0] if acc is None else self.ctx.bf_or([acc, part[0]], num_bits)
        if acc is None:
            return 0
        return int(self.ctx.bf_popcount([acc], num_bits)[0])

    # ------------------------------------------------------------------ bfoperate --or
    def union(self, src_paths: Sequence[str], out_path: str) -> None:
        rows = [self.filter(p) for p in src_paths]
        out = self.ctx.bf_or(rows, self.num_bits)
        if self.stride > self.words:
            out[self.words:] = 0
        self._write(out_path, out)
        self._put(out_path, out)

    def copy(self, src_path: str, out_path: str) -> None:
        t = self.filter(src_path).clone()
        self._write(out_path, t)
        self._put(out_path, t)

    # ------------------------------------------------------------------ dumpbf --show:density
    def popcounts(self, paths: Sequence[str]) -> List[int]:
        if not paths:
            return []
        rows = [self.filter(p) for p in paths]
        return [int(x) for x in self.ctx.bf_popcount(rows, self.num_bits).cpu()]

    # ------------------------------------------------------------------ bfdistance
    def dist_one_vs_many(self, focus_path: str, target_paths: Sequence[str]) -> List[Tuple[int, int]]:
        """[(intersection, union)] of the focus against every target, in target order."""
        if not target_paths:
            return []
        inter, union = self.ctx.bf_dist(self.filter(focus_path), [self.filter(p) for p in target_paths], self.num_bits)
        i, u = inter.cpu().tolist(), union.cpu().tolist()
        return list(zip(i, u))

    def dist_all_pairs(self, paths: Sequence[str]) -> np.ndarray:
        """Dense intersection matrix; diagonal = popcounts (union_ij = I_ii + I_jj - I_ij)."""
        rows = [self.filter(p) for p in paths]
        return self.ctx.bf_allpairs(rows, self.num_bits).cpu().numpy()

    def dist_all_pairs_many(self, groups: Sequence[Sequence[str]]) -> List[np.ndarray]:
        """dist_all_pairs for many clusters: every all-pairs launch is queued first, the matrices come back in ONE
        device->host copy (one synchronisation for the whole batch instead of one per cluster)."""
        mats = [self.ctx.bf_allpairs([self.filter(p) for p in g], self.num_bits) for g in groups]
        if not mats:
            return []
        flat = torch.cat([m.reshape(-1) for m in mats]).cpu().numpy()
        out, off = [], 0
        for g in groups:
            n = len(g)
            out.append(flat[off: off + n * n].reshape(n, n))
            off += n * n
        return out

    # ------------------------------------------------------------------ query --distinctkmers
    def query_positions(self, fasta_path: str) -> torch.Tensor:
        """Distinct hash positions of ALL k-mers of the file (no min filter), ascending, on the device.
        They are the set bits of the file's own min=1 filter (App. A6), so: K1 + ordered extraction."""
        q = self.sketch_resident([fasta_path], min_abund=1)
        return self.ctx.extract_positions(q[0], self.num_bits)

    def query_hits(self, positions: torch.Tensor, leaf_paths: Sequence[str]) -> Tuple[List[int], int]:
        """(hits per leaf, total) for one query against the leaves of a flat tree."""
        total = int(positions.numel())
        if not leaf_paths:
            return [], total
        sliced = self._sliced_tree(leaf_paths)
        if sliced is not None:
            self.sliced_probes += 1
            hits = self.ctx.query_batch_sliced(sliced, len(leaf_paths), 0, self.num_bits, [positions])
        else:
            hits = self.ctx.query_batch([self.filter(p) for p in leaf_paths], [positions])
        return [int(x) for x in hits[0].cpu()], total

    def _sliced_tree(self, leaf_paths: Sequence[str]) -> Optional[torch.Tensor]:
        """Bit-sliced matrix of a flat tree's leaves, or None when the row-major probe should serve the query.
        The row-major probe reads one 32-byte sector per (position, leaf); the bit-sliced one reads 1/8 byte per
        (position, leaf) but needs the leaves transposed first (one pass over all of them).  A tree with many leaves
        that is queried again -- `profile` walks the same trees for every input genome -- gets transposed on its second
        query and the matrix is kept (LRU, bounded in bytes; dropped when one of its filters is rewritten)."""
        n = len(leaf_paths)
        if n < self.sliced_min_leaves:
            return None
        key = tuple(os.path.abspath(p) for p in leaf_paths)
        t = self._sliced.get(key)
        if t is not None:
            self._sliced.move_to_end(key)
            return t
        seen = self._tree_queries.get(key, 0) + 1
        self._tree_queries[key] = seen
        nbytes = self.num_bits * self.ctx.sliced_row_words(n) * 4
        if seen < 2 or nbytes > self.sliced_cache_bytes:
            return None
        t = self.ctx.bitslice([self.filter(p) for p in key], 0, self.num_bits)
        self._sliced[key] = t
        self._sliced_bytes += nbytes
        while self._sliced_bytes > self.sliced_cache_bytes and len(self._sliced) > 1:
            _, old = self._sliced.popitem(last=False)
            self._sliced_bytes -= old.numel() * 4
        return t

    def query_many(self, query_filters: Sequence[torch.Tensor], leaf_paths: Sequence[str]):
        """hits[q][l] and totals[q] for device-resident query filters against the given leaves."""
        leaves = [self.filter(p) for p in leaf_paths]
        positions = self.ctx.extract_positions_batch(list(query_filters), self.num_bits)
        hits = self.ctx.query_batch(leaves, positions)
        totals = [int(p.numel()) for p in positions]
        return hits.cpu().numpy(), totals
