"""The product on several GPUs of one box: `MultiGpuBackend` has the interface of backend.CudaBackend and spreads every
operator over one CudaBackend (one msbt_ctx, one filter cache) per device, driven by one host thread per device -- no
forks after CUDA initialisation.  It replaces the reference's process fan-outs (metasbt.py:542-543, 1124-1143, 1231-1242;
core.py:4036-4060) the way SURVEY.md 8e partitions the path:

* makebf      genomes are independent: contiguous blocks balanced by size (shard.genome_blocks), one block per device, no
              exchange.  A filter's HOME is the device that built (or first loaded) it.
* bfoperate   every device ORs the children it is home to (K2); the partial unions meet on one device over NVLink
              (peer copies) and are OR-ed there.
* dumpbf / bfdistance   operands are processed where they live; a focus filter is copied to the devices that hold its
              targets; an all-pairs cluster is gathered on the device that already holds most of it.
* query       MAGs are hashed in blocks, one block per device (K1 + extraction).  Small flat trees are probed row-major on
              the device that holds the MAG's positions.  A LARGE flat tree keeps its bit-sliced leaf matrix sharded by
              contiguous bit range (shard.bit_ranges): device r owns rows [r*m/P, (r+1)*m/P) of every leaf; each device
              extracts, per destination, the positions of its MAGs inside that range and sends them there (peer copies),
              every device probes its shard with the positions of ALL MAGs, and the partial hit matrices are summed with
              one NCCL reduce (torch.cuda.nccl: single process, one communicator per device).

Select it with MSBT_DEVICES="0,1,2,3" (or "all"); core.get_backend then returns a MultiGpuBackend instead of a CudaBackend.
There is no CPU implementation here either.
"""

from __future__ import annotations

import os
from collections import OrderedDict
from concurrent.futures import ThreadPoolExecutor
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import shard
from .backend import CudaBackend, QueryBatch


def devices_from_env() -> Optional[List[int]]:
    """MSBT_DEVICES: "all" or a comma-separated list of CUDA device indices; unset / one device -> None (single GPU)."""
    text = os.environ.get("MSBT_DEVICES", "").strip()
    if not text:
        return None
    if text == "all":
        devs = list(range(torch.cuda.device_count()))
    else:
        devs = [int(t) for t in text.split(",") if t.strip()]
    return devs if len(devs) > 1 else None


class MultiQueryBatch:
    """Query positions of a batch hashed in blocks over the devices: part[d] is device d's QueryBatch; query i of the batch
    is query `local[i]` of part[`where[i]`]."""

    def __init__(self, parts: List[Optional[QueryBatch]], where: List[int], local: List[int]):
        self.parts, self.where, self.local = parts, where, local

    def __len__(self) -> int:
        return len(self.where)


class _ShardedTree:
    """A large flat tree: device r holds rows [rb_r, re_r) of the bit-sliced matrix of all its leaves."""

    def __init__(self, n_leaves: int, sliced: List[torch.Tensor], ranges: List[Tuple[int, int]]):
        self.n_leaves, self.sliced, self.ranges = n_leaves, sliced, ranges
        self.nbytes = sum(t.numel() * 4 for t in sliced)


class MultiGpuBackend:
    def __init__(self, kmer_size: int, num_bits: int, devices: Sequence[int], **backend_kwargs):
        if len(devices) < 2:
            raise ValueError("MultiGpuBackend needs at least two devices (use backend.CudaBackend for one)")
        if not torch.cuda.is_available() or max(devices) >= torch.cuda.device_count():
            raise RuntimeError("metasbt_b200 needs the CUDA devices it was asked for: the hot path has no CPU fallback")
        self.k, self.num_bits = int(kmer_size), int(num_bits)
        self.devices = [int(d) for d in devices]
        self.P = len(self.devices)
        self.parts = [CudaBackend(kmer_size, num_bits, device=d, **backend_kwargs) for d in self.devices]
        self.words, self.stride, self.row_bytes = self.parts[0].words, self.parts[0].stride, self.parts[0].row_bytes
        self.position_bytes = sum(p.position_bytes for p in self.parts)
        self._pool = ThreadPoolExecutor(max_workers=self.P)
        self._home: Dict[str, int] = dict()
        self._rr = 0
        self.sliced_min_leaves = 256
        self._trees: "OrderedDict[Tuple[str, ...], _ShardedTree]" = OrderedDict()
        self._tree_queries: "OrderedDict[Tuple[str, ...], int]" = OrderedDict()
        self.sliced_probes = 0
        self.nccl_reduces = 0

    # ------------------------------------------------------------------ plumbing
    def _each(self, jobs: Dict[int, Callable[[], object]]) -> Dict[int, object]:
        """Run one callable per device index, each on that device's thread; exceptions propagate."""
        futures = {d: self._pool.submit(fn) for d, fn in jobs.items()}
        return {d: f.result() for d, f in futures.items()}

    def _home_of(self, path: str) -> int:
        path = os.path.abspath(path)
        d = self._home.get(path)
        if d is None:
            d = self._rr % self.P  # first seen as a file on disk: spread over the devices
            self._rr += 1
            self._home[path] = d
        return d

    def _set_home(self, path: str, d: int) -> None:
        """`path` was (re)written on device d: that copy is the filter, replicas elsewhere and sharded trees are stale."""
        path = os.path.abspath(path)
        self._home[path] = d
        for i, part in enumerate(self.parts):
            if i != d:
                part.forget(path)
        for key in [k for k in self._trees if path in k]:
            del self._trees[key]
        for key in [k for k in self._tree_queries if path in k]:
            del self._tree_queries[key]

    def _ensure_on(self, d: int, path: str) -> None:
        """Make the filter resident in device d's cache: from its home device over NVLink (peer copy) when it is cached
        there, from the file otherwise.  Called from the coordinating thread only."""
        path = os.path.abspath(path)
        part = self.parts[d]
        if path in part._cache:
            return
        h = self._home_of(path)
        src = self.parts[h]._cache.get(path) if h != d else None
        if src is not None:
            part._put(path, src.to(part.ctx.device))
        else:
            part.filter(path)

    def filter(self, path: str) -> torch.Tensor:
        d = self._home_of(path)
        return self.parts[d].filter(path)

    def forget(self, path: str) -> None:
        path = os.path.abspath(path)
        for part in self.parts:
            part.forget(path)
        self._home.pop(path, None)

    # ------------------------------------------------------------------ makebf
    def sketch_many(self, fasta_paths: Sequence[str], out_paths: Sequence[str], min_abund: int) -> None:
        if not fasta_paths:
            return
        blocks = shard.genome_blocks([os.path.getsize(p) for p in fasta_paths], self.P, self.k)
        jobs = {}
        for d, (b, e) in enumerate(blocks):
            if e > b:
                jobs[d] = (lambda d=d, b=b, e=e: self.parts[d].sketch_many(fasta_paths[b:e], out_paths[b:e], min_abund))
        self._each(jobs)
        for d, (b, e) in enumerate(blocks):
            for out in out_paths[b:e]:
                self._set_home(out, d)

    def count_union_bits(self, fasta_paths: Sequence[str], kmer_size: int, num_bits: int, modulus: int) -> int:
        # one shared counting filter: a single pass on one device (estimate.py; not worth an exchange)
        return self.parts[0].count_union_bits(fasta_paths, kmer_size, num_bits, modulus)

    # k-mer size search (kopt.py): at most kopt.MAX_GENOMES small filters per k -- one device
    @property
    def device(self):
        return self.parts[0].device

    def read_fasta(self, path: str) -> bytes:
        return self.parts[0].read_fasta(path)

    def kmer_set_profile(self, fasta_paths: Sequence[str], kmer_size: int, num_bits: int, modulus: int):
        return self.parts[0].kmer_set_profile(fasta_paths, kmer_size, num_bits, modulus)

    def kmer_set_profile_done(self) -> None:
        self.parts[0].kmer_set_profile_done()

    # ------------------------------------------------------------------ bfoperate --or
    def union(self, src_paths: Sequence[str], out_path: str) -> None:
        by_dev: Dict[int, List[str]] = dict()
        for p in src_paths:
            by_dev.setdefault(self._home_of(p), []).append(p)
        target = max(by_dev, key=lambda d: len(by_dev[d]))

        def partial(d: int):
            part = self.parts[d]
            acc = None
            step = part._chunk()
            for lo in range(0, len(by_dev[d]), step):
                rows = [part.filter(p) for p in by_dev[d][lo: lo + step]]
                acc = part.ctx.bf_or(rows if acc is None else [acc] + rows, self.num_bits)
            return acc

        partials = self._each({d: (lambda d=d: partial(d)) for d in by_dev})
        tpart = self.parts[target]
        rows = [partials[target]] + [partials[d].to(tpart.ctx.device) for d in by_dev if d != target]  # NVLink peer copies
        out = tpart.ctx.bf_or(rows, self.num_bits) if len(rows) > 1 else rows[0]
        if self.stride > self.words:
            out[self.words:] = 0
        tpart._write(out_path, out)
        tpart._put(out_path, out)
        self._set_home(out_path, target)

    def union_many(self, jobs: Sequence[Tuple[Sequence[str], str]]) -> None:
        """All clusters of one level: a job whose sources live on one device runs there (jobs of different devices in
        parallel, each through CudaBackend.union_many); a job that spans devices goes through union()."""
        local: Dict[int, List[Tuple[Sequence[str], str]]] = dict()
        spanning = []
        for srcs, out in jobs:
            homes = {self._home_of(p) for p in srcs}
            if len(homes) == 1:
                local.setdefault(homes.pop(), []).append((srcs, out))
            else:
                spanning.append((srcs, out))
        self._each({d: (lambda d=d: self.parts[d].union_many(local[d])) for d in local})
        for d, js in local.items():
            for _srcs, out in js:
                self._set_home(out, d)
        for srcs, out in spanning:
            self.union(srcs, out)

    def copy(self, src_path: str, out_path: str) -> None:
        d = self._home_of(src_path)
        self.parts[d].copy(src_path, out_path)
        self._set_home(out_path, d)

    # ------------------------------------------------------------------ dumpbf --show:density
    def popcounts(self, paths: Sequence[str]) -> List[int]:
        by_dev: Dict[int, List[int]] = dict()
        for i, p in enumerate(paths):
            by_dev.setdefault(self._home_of(p), []).append(i)
        got = self._each({d: (lambda d=d: self.parts[d].popcounts([paths[i] for i in by_dev[d]])) for d in by_dev})
        out = [0] * len(paths)
        for d, idxs in by_dev.items():
            for i, c in zip(idxs, got[d]):
                out[i] = c
        return out

    # ------------------------------------------------------------------ bfdistance
    def dist_one_vs_many(self, focus_path: str, target_paths: Sequence[str]) -> List[Tuple[int, int]]:
        return self.dist_pairs([(focus_path, t) for t in target_paths])

    def dist_pairs(self, pairs: Sequence[Tuple[str, str]]) -> List[Tuple[int, int]]:
        """Every pair is computed on the home device of its SECOND member (the database side of profile_many's requests);
        first members that live elsewhere are copied over once."""
        if not pairs:
            return []
        by_dev: Dict[int, List[int]] = dict()
        for i, (_a, b) in enumerate(pairs):
            by_dev.setdefault(self._home_of(b), []).append(i)
        for d, idxs in by_dev.items():
            for a in OrderedDict((pairs[i][0], None) for i in idxs):
                self._ensure_on(d, a)
        got = self._each({d: (lambda d=d: self.parts[d].dist_pairs([pairs[i] for i in by_dev[d]])) for d in by_dev})
        out: List[Tuple[int, int]] = [(0, 0)] * len(pairs)
        for d, idxs in by_dev.items():
            for i, c in zip(idxs, got[d]):
                out[i] = c
        return out

    def _gather_group(self, paths: Sequence[str]) -> int:
        """Device that already holds most of the group; the rest is copied there."""
        votes: Dict[int, int] = dict()
        for p in paths:
            votes[self._home_of(p)] = votes.get(self._home_of(p), 0) + 1
        d = max(votes, key=lambda x: votes[x])
        if len(paths) <= self.parts[d]._chunk(2):
            for p in paths:
                self._ensure_on(d, p)
        return d

    def dist_all_pairs(self, paths: Sequence[str]) -> np.ndarray:
        d = self._gather_group(paths)
        return self.parts[d].dist_all_pairs(paths)

    def dist_all_pairs_many(self, groups: Sequence[Sequence[str]]) -> List[np.ndarray]:
        where = [self._gather_group(g) for g in groups]
        by_dev: Dict[int, List[int]] = dict()
        for i, d in enumerate(where):
            by_dev.setdefault(d, []).append(i)
        got = self._each({d: (lambda d=d: self.parts[d].dist_all_pairs_many([groups[i] for i in by_dev[d]])) for d in by_dev})
        out: List[Optional[np.ndarray]] = [None] * len(groups)
        for d, idxs in by_dev.items():
            for i, m in zip(idxs, got[d]):
                out[i] = m
        return out  # type: ignore[return-value]

    # ------------------------------------------------------------------ query --distinctkmers
    def query_positions_many(self, fasta_paths: Sequence[str]) -> MultiQueryBatch:
        blocks = shard.genome_blocks([os.path.getsize(p) for p in fasta_paths], self.P, self.k)
        got = self._each({d: (lambda d=d, b=b, e=e: self.parts[d].query_positions_many(fasta_paths[b:e]))
                          for d, (b, e) in enumerate(blocks) if e > b})
        parts: List[Optional[QueryBatch]] = [got.get(d) for d in range(self.P)]
        where, local = [], []
        for d, (b, e) in enumerate(blocks):
            where.extend([d] * (e - b))
            local.extend(range(e - b))
        return MultiQueryBatch(parts, where, local)

    def query_positions(self, fasta_path: str) -> MultiQueryBatch:
        return self.query_positions_many([fasta_path])

    def query_totals(self, qb: MultiQueryBatch) -> List[int]:
        return [qb.parts[d].total(l) for d, l in zip(qb.where, qb.local)]

    def query_hits_many(self, qb: MultiQueryBatch, idxs: Sequence[int], leaf_paths: Sequence[str]):
        """-> a pending object for fetch_hits: hit rows of queries `idxs` against the leaves of one flat tree."""
        nq, nl = len(idxs), len(leaf_paths)
        if nq == 0 or nl == 0:
            return ("rows", [], np.zeros((nq, nl), dtype=np.int32))
        n_pos = sum(qb.parts[qb.where[i]].total(qb.local[i]) for i in idxs)
        tree = self._sharded_tree(leaf_paths, n_pos)
        if tree is not None:
            self.sliced_probes += nq
            return ("device", self._probe_sharded(tree, qb, idxs))
        # small tree: row-major probe on the device that holds each query's positions, leaves replicated there
        by_dev: Dict[int, List[int]] = dict()
        for r, i in enumerate(idxs):
            by_dev.setdefault(qb.where[i], []).append(r)
        for d in by_dev:
            for p in leaf_paths:
                self._ensure_on(d, p)
        got = self._each({d: (lambda d=d: self.parts[d].query_hits_many(qb.parts[d], [qb.local[idxs[r]] for r in by_dev[d]], leaf_paths))
                          for d in by_dev})
        return ("rows", [(by_dev[d], got[d]) for d in by_dev], (nq, nl))

    def fetch_hits(self, pending) -> List[np.ndarray]:
        """One device->host copy per device for all pending hit matrices of a tree level."""
        per_dev: Dict[int, List[torch.Tensor]] = dict()
        for item in pending:
            if item[0] == "device":
                per_dev.setdefault(item[1].device.index, []).append(item[1])
            elif item[0] == "rows" and isinstance(item[1], list):
                for _rows, t in item[1]:
                    per_dev.setdefault(t.device.index, []).append(t)
        host: Dict[int, np.ndarray] = {d: torch.cat([t.reshape(-1) for t in ts]).cpu().numpy() for d, ts in per_dev.items()}
        cursor = {d: 0 for d in per_dev}

        def take(t: torch.Tensor) -> np.ndarray:
            d = t.device.index
            out = host[d][cursor[d]: cursor[d] + t.numel()].reshape(t.shape)
            cursor[d] += t.numel()
            return out

        out = []
        for item in pending:
            if item[0] == "device":
                out.append(take(item[1]))
            elif isinstance(item[1], list) and isinstance(item[2], tuple):
                full = np.zeros(item[2], dtype=np.int32)
                for rows, t in item[1]:
                    full[rows] = take(t)
                out.append(full)
            else:
                out.append(item[2])
        return out

    def query_hits(self, positions: MultiQueryBatch, leaf_paths: Sequence[str]) -> Tuple[List[int], int]:
        total = self.query_totals(positions)[0]
        if not leaf_paths:
            return [], total
        hits = self.fetch_hits([self.query_hits_many(positions, [0], leaf_paths)])[0]
        return [int(x) for x in hits[0]], total

    # ------------------------------------------------------------------ large trees: bit-range sharded sliced matrix
    def _sharded_tree(self, leaf_paths: Sequence[str], n_positions: int) -> Optional[_ShardedTree]:
        n = len(leaf_paths)
        if n < self.sliced_min_leaves:
            return None
        key = tuple(os.path.abspath(p) for p in leaf_paths)
        tree = self._trees.get(key)
        if tree is not None:
            self._trees.move_to_end(key)
            return tree
        seen = self._tree_queries.pop(key, 0) + 1
        self._tree_queries[key] = seen
        while len(self._tree_queries) > 4096:
            self._tree_queries.popitem(last=False)
        row_words = self.parts[0].ctx.sliced_row_words(n)
        nbytes = self.num_bits * row_words * 4
        if nbytes // self.P > min(p.sliced_cache_bytes for p in self.parts):
            return None
        rowmajor = n_positions * n * 32
        sliced_cost = 2 * n * self.row_bytes + n_positions * row_words * 4
        if seen < 2 and rowmajor <= sliced_cost:
            return None
        ranges = shard.bit_ranges(self.num_bits, self.P)
        for p in key:  # leaves resident on their home devices (loaded from file if need be)
            self.parts[self._home_of(p)].filter(p)

        def build(d: int) -> torch.Tensor:
            part = self.parts[d]
            rb, re_ = ranges[d]
            w0, w1 = rb // 64, (re_ + 63) // 64
            out = torch.empty((re_ - rb, row_words), dtype=torch.int32, device=part.ctx.device)
            group = max(32, (part._chunk() * self.P) // 32 * 32)  # leaves per pass: their RANGE slices fit the chunk budget
            for g0 in range(0, n, group):
                g1 = min(n, g0 + group)
                # this device's bit range of every leaf of the group, fetched from the leaf's home device (peer copies)
                rows = torch.stack([self.parts[self._home[p]]._cache[p][w0:w1].to(part.ctx.device, non_blocking=True)
                                    if p in self.parts[self._home[p]]._cache else self.parts[self._home[p]].filter(p)[w0:w1].to(part.ctx.device)
                                    for p in key[g0:g1]])
                sub = part.ctx.bitslice([rows[i] for i in range(g1 - g0)], 0, re_ - rb)
                out[:, g0 // 32: g0 // 32 + sub.shape[1]] = sub
            return out

        # built one device after the other from the coordinating thread: the source caches are read while nothing mutates them
        sliced = [build(d) for d in range(self.P)]
        tree = _ShardedTree(n, sliced, ranges)
        self._trees[key] = tree
        budget = sum(p.sliced_cache_bytes for p in self.parts)
        while sum(t.nbytes for t in self._trees.values()) > budget and len(self._trees) > 1:
            self._trees.popitem(last=False)
        return tree

    def _probe_sharded(self, tree: _ShardedTree, qb: MultiQueryBatch, idxs: Sequence[int]) -> torch.Tensor:
        """hits[len(idxs), L] on the first device: per-destination extraction on the hashing devices, peer copies of the
        position lists, one probe per shard, NCCL reduce of the partial matrices."""
        P = self.P
        by_src: Dict[int, List[int]] = dict()  # source device -> rows of `idxs` hashed there
        for r, i in enumerate(idxs):
            by_src.setdefault(qb.where[i], []).append(r)

        def split(s: int):
            """On source device s: for every destination range, the positions of its queries inside that range (relative
            to the range start), packed, with per-query counts."""
            part, q = self.parts[s], qb.parts[s]
            pos = q.positions
            out = []
            bounds = torch.tensor([rb for rb, _ in tree.ranges] + [self.num_bits], dtype=torch.int64, device=pos.device)
            for r in by_src[s]:
                b, e = q.offsets[qb.local[idxs[r]]], q.offsets[qb.local[idxs[r]] + 1]
                seg = pos[b:e]
                # positions are ascending uint32 bit patterns in an int32 tensor: compare as int64
                cut = torch.searchsorted(seg.to(torch.int64) & 0xFFFFFFFF, bounds)
                out.append((seg, cut))
            cuts = torch.stack([c for _, c in out]).cpu().tolist() if out else []
            return [(seg, c) for (seg, _), c in zip(out, cuts)]

        pieces = self._each({s: (lambda s=s: split(s)) for s in by_src})
        # destination d: [for every row r of idxs: positions inside range d, relative], gathered by peer copies
        recv: List[List[Optional[torch.Tensor]]] = [[None] * len(idxs) for _ in range(P)]
        for s, rows in by_src.items():
            for r, (seg, cut) in zip(rows, pieces[s]):
                for d in range(P):
                    part = seg[cut[d]: cut[d + 1]]
                    # uint32 bit patterns in an int32 tensor: rebase in 64 bits (a range start can be 2^31 and more)
                    rel = ((part.to(torch.int64) & 0xFFFFFFFF) - tree.ranges[d][0]).to(torch.int32)
                    recv[d][r] = rel.to(self.parts[d].ctx.device, non_blocking=True)

        def probe(d: int) -> torch.Tensor:
            part = self.parts[d]
            lists = recv[d]
            sizes = [int(t.numel()) for t in lists]
            allpos = torch.cat(lists) if sum(sizes) else torch.zeros(1, dtype=torch.int32, device=part.ctx.device)
            ranges, cur = [], 0
            for n_ in sizes:
                ranges.append((cur, cur + n_))
                cur += n_
            rb, re_ = tree.ranges[d]
            return part.ctx.query_sliced_ranges(tree.sliced[d], tree.n_leaves, 0, re_ - rb, allpos, ranges)

        partial = self._each({d: (lambda d=d: probe(d)) for d in range(P)})
        parts = [partial[d] for d in range(P)]
        for d in range(P):
            torch.cuda.synchronize(self.parts[d].ctx.device)
        import torch.cuda.nccl as nccl
        nccl.reduce(parts, output=parts[0], root=0)  # sum of the partial hit matrices over NVLink
        self.nccl_reduces += 1
        return parts[0]
