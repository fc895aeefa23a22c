"""File-level operators: the five `howdesbt` subcommands MetaSBT spawns, as in-process batch calls
over HBM-resident filters.  core.py talks to this interface only.

    makebf      -> CudaBackend.sketch_many      (core.py:3920-3931)
    bfoperate   -> CudaBackend.union            (core.py:3851-3857)
    dumpbf      -> CudaBackend.popcounts        (core.py:3588-3593)
    bfdistance  -> CudaBackend.dist_one_vs_many / dist_pairs / dist_all_pairs   (core.py:1747-1774, 4036-4081)
    query       -> CudaBackend.query_positions_many + query_hits_many           (core.py:2039-2047)

Filters are cached on the device by absolute path (LRU, bounded in bytes), so that the repeated
reads the reference pays per spawn (every `bfdistance`/`query` call re-reads every listed file)
happen once.  `.bf` files stay the persistent store and are byte-compatible (bffile.py).
Every operator walks its operands in chunks that fit the cache budget, so a database larger than the
card's memory streams through it instead of exhausting it (the reference streams files per spawn).
Budgets come from cudaMemGetInfo at construction.
There is no CPU implementation here: without libmsbt.so and a CUDA device, construction fails.
"""

from __future__ import annotations

import os
from collections import OrderedDict
from concurrent.futures import ThreadPoolExecutor
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import bffile, ops


def read_fasta_bytes(path: str) -> bytes:
    with open(path, "rb") as f:
        return f.read()


def read_files(paths: Sequence[str], threads: Optional[int] = None) -> List[bytes]:
    """Read many FASTA files on a thread pool (file reads release the GIL)."""
    threads = threads or min(32, os.cpu_count() or 1)
    if len(paths) <= 1 or threads <= 1:
        return [read_fasta_bytes(p) for p in paths]
    with ThreadPoolExecutor(max_workers=threads) as ex:
        return list(ex.map(read_fasta_bytes, paths))


# All pairs through K3 (AND + popcount over whole filters: n^2 / 2 filter passes) or through K4 (transpose once, then every
# filter's set bits probe the bit-sliced matrix: work ~ set bits x row width).  Rates measured on B200:
# the carry-save all-pairs kernel does 591 k pairs/s at 2^28 bits; the row-coherent probe 24 G positions/s on 128-byte rows
# and 8.6 TB/s of row bytes on wide rows; the transposition moves 2 x n x m / 8 bytes at 3.2 TB/s per pass.
PROBE_MIN_FILTERS = 192
BIT_ROUTE_PAIRS_PER_S_AT_2P28 = 591e3
PROBE_POSITIONS_PER_S = 24e9
PROBE_ROW_BYTES_PER_S = 8.6e12
TRANSPOSE_BYTES_PER_S = 3.2e12


def allpairs_route(n: int, num_bits: int, total_popcount: int, row_words: int, slice_bytes: int) -> str:
    """'probe' when the K4 formulation of an n-filter all-pairs matrix is expected to be clearly faster than K3's."""
    if n < 2 or num_bits > (1 << 32):
        return "bit"
    t_bit = (n * (n + 1) / 2.0) * (num_bits / float(1 << 28)) / BIT_ROUTE_PAIRS_PER_S_AT_2P28
    row_bytes = row_words * 4
    passes = max(1, -(-(num_bits * row_bytes) // max(int(slice_bytes), row_bytes)))
    t_probe = total_popcount / min(PROBE_POSITIONS_PER_S, PROBE_ROW_BYTES_PER_S / row_bytes)
    t_probe += 2.0 * n * (num_bits / 8.0) / TRANSPOSE_BYTES_PER_S          # the transposition, once over all passes
    t_probe += passes * n * (num_bits / 8.0) / 2.7e12                        # one extraction of every filter per pass
    return "probe" if t_probe < 0.7 * t_bit else "bit"


def _retry_after_oom(method):
    """An operator that runs out of device memory drops the backend's caches (row-major filters, bit-sliced matrices,
    torch's cached blocks) and runs once more -- budgets come from cudaMemGetInfo at construction, other users of the
    device may have grown since."""
    import functools

    @functools.wraps(method)
    def wrapper(self, *args, **kwargs):
        try:
            return method(self, *args, **kwargs)
        except torch.cuda.OutOfMemoryError:
            self.release_memory()
            return method(self, *args, **kwargs)
    return wrapper


class QueryBatch:
    """Distinct hash positions of a batch of queries on the device: `positions` (int32 tensor holding uint32 bit
    patterns) carries every query's ascending positions back to back, query i owns [offsets[i], offsets[i + 1])."""

    def __init__(self, positions: torch.Tensor, offsets: Sequence[int]):
        self.positions = positions
        self.offsets = [int(x) for x in offsets]

    def __len__(self) -> int:
        return len(self.offsets) - 1

    def total(self, i: int) -> int:
        return self.offsets[i + 1] - self.offsets[i]

    def ranges(self, idxs: Sequence[int]) -> List[Tuple[int, int]]:
        return [(self.offsets[i], self.offsets[i + 1]) for i in idxs]


class CudaBackend:
    # fractions of the device memory that is free at construction
    CACHE_FRACTION = 0.45          # row-major filters cached by path
    SLICED_FRACTION = 0.25         # bit-sliced leaf matrices of large flat trees
    BATCH_FRACTION = 0.10          # the filter matrix one makebf / query batch is built into

    def __init__(self, kmer_size: int, num_bits: int, device: Optional[int] = None,
                 cache_bytes: Optional[int] = None, sketch_batch_bytes: Optional[int] = None,
                 sliced_cache_bytes: Optional[int] = None):
        self.k = int(kmer_size)
        self.num_bits = int(num_bits)
        self.ctx = ops.default_context(device)
        self.words = ops.filter_words(self.num_bits)
        self.stride = ops.filter_stride_words(self.num_bits)
        self.row_bytes = self.stride * 8
        free, _total = torch.cuda.mem_get_info(self.ctx.device)
        self.cache_bytes = int(free * self.CACHE_FRACTION) if cache_bytes is None else int(cache_bytes)
        self.sliced_cache_bytes = int(free * self.SLICED_FRACTION) if sliced_cache_bytes is None else int(sliced_cache_bytes)
        batch_bytes = int(free * self.BATCH_FRACTION) if sketch_batch_bytes is None else int(sketch_batch_bytes)
        self.sketch_batch = max(1, batch_bytes // self.row_bytes)
        self.position_bytes = batch_bytes  # packed query positions held at once by profile_many
        self._cache: "OrderedDict[str, torch.Tensor]" = OrderedDict()
        self._bytes = 0
        # bit-sliced leaf matrices of flat trees that are queried repeatedly (see _sliced_tree)
        self.sliced_min_leaves = 256
        self._sliced: "OrderedDict[Tuple[str, ...], torch.Tensor]" = OrderedDict()
        self._sliced_bytes = 0
        self._tree_queries: "OrderedDict[Tuple[str, ...], int]" = OrderedDict()
        self.sliced_probes = 0  # queries served by the bit-sliced probe (observability / tests)
        self._pipes: Dict[int, object] = dict()       # makebf pipelines by min_abund (pinned slots allocated once)
        self._text: Optional[torch.Tensor] = None     # pinned FASTA staging buffer of sketch_many

    # ------------------------------------------------------------------ device filter cache
    def _put(self, path: str, t: torch.Tensor) -> None:
        path = os.path.abspath(path)
        # the filter stored at `path` is (re)written: every bit-sliced matrix that contains it is stale, whether or not
        # the row-major copy is still cached
        self._drop_sliced(path)
        if path in self._cache:
            self._bytes -= self._cache.pop(path).numel() * 8
        self._cache[path] = t
        self._bytes += t.numel() * 8
        while self._bytes > self.cache_bytes and len(self._cache) > 1:
            _, old = self._cache.popitem(last=False)
            self._bytes -= old.numel() * 8

    def release_memory(self) -> None:
        """Drop every cached filter and bit-sliced matrix and return torch's cached blocks to the driver."""
        self._cache.clear()
        self._bytes = 0
        self._sliced.clear()
        self._sliced_bytes = 0
        self._pipes.clear()
        self._set_text_key = self._set_text = None
        torch.cuda.empty_cache()

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
        for key in [k for k in self._tree_queries if path in k]:
            del self._tree_queries[key]

    def filter(self, path: str) -> torch.Tensor:
        """Device tensor (int64[stride]) of the filter stored at `path` (cached)."""
        t = self._cache.get(path)  # callers nearly always pass the absolute path the cache is keyed by
        if t is None:
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

    def _chunk(self, minimum: int = 1) -> int:
        """How many filters an operator keeps alive at once: half the cache budget (the cache itself may hold the other
        half), so that operand lists larger than the card stream through instead of exhausting it."""
        return max(minimum, (self.cache_bytes // 2) // self.row_bytes)

    def _write(self, path: str, t: torch.Tensor) -> None:
        header = bffile.BloomHeader(kmer_size=self.k, num_bits=self.num_bits)
        bffile.write_bf(path, header, t[: self.words].cpu().numpy().view(np.uint64))

    # ------------------------------------------------------------------ makebf
    def _pipeline(self, min_abund: int):
        from . import pipeline
        pipe = self._pipes.get(min_abund)
        if pipe is None:
            chunk = int(min(max((2 << 30) // self.row_bytes, 8), 1024, max(self.sketch_batch // 2, 1)))
            pipe = pipeline.SketchPipeline(self.ctx, self.k, self.num_bits, min_abund=min_abund, chunk_genomes=chunk, slots=2)
            self._pipes[min_abund] = pipe
        return pipe

    @_retry_after_oom
    def sketch_many(self, fasta_paths: Sequence[str], out_paths: Sequence[str], min_abund: int) -> None:
        """howdesbt makebf --k --min --bits --hashes=1 --seed=0,0 <fasta> --out=<bf> for every pair.
        The files of a group are read into one pinned text buffer and go through pipeline.SketchPipeline: raw FASTA text
        host -> device, parsing and 2-bit packing on the GPU (msbt_fasta_pack_device), K1, filters back into pinned slots
        (as position lists when that is the smaller thing to move) while the next chunk is being built, `.bf` files
        written from the slots by a thread pool.  Fresh filters also stay cached on the device (index ORs them next)."""
        if not fasta_paths:
            return
        pipe = self._pipeline(min_abund)
        header = bffile.BloomHeader(kmer_size=self.k, num_bits=self.num_bits)
        group_bytes = 1 << 30
        lo = 0
        with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as writers:
            while lo < len(fasta_paths):
                hi, size = lo, 0
                while hi < len(fasta_paths) and (hi == lo or size + os.path.getsize(fasta_paths[hi]) <= group_bytes):
                    size += os.path.getsize(fasta_paths[hi])
                    hi += 1
                texts = read_files(fasta_paths[lo:hi])
                offsets = [0]
                for t in texts:
                    offsets.append(offsets[-1] + len(t))
                if self._text is None or self._text.numel() < max(offsets[-1], 1):
                    self._text = torch.empty(max(offsets[-1], 1), dtype=torch.uint8, pin_memory=True)
                hv = self._text.numpy()
                for i, t in enumerate(texts):
                    if len(t):
                        hv[offsets[i]: offsets[i + 1]] = np.frombuffer(t, dtype=np.uint8)
                del texts
                outs = out_paths[lo:hi]

                def sink(first: int, words: np.ndarray, outs=outs):
                    jobs = [writers.submit(bffile.write_bf, outs[first + i], header, words[i]) for i in range(len(words))]
                    for j in jobs:
                        j.result()

                def on_device(first: int, filters: torch.Tensor, outs=outs):
                    for i in range(filters.shape[0]):
                        self._put(outs[first + i], filters[i].clone())

                pipe.run(self._text, offsets, sink=sink, on_device=on_device)
                lo = hi

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

    # ------------------------------------------------------------------ k-mer size search (kopt.py)
    @property
    def device(self) -> torch.device:
        return self.ctx.device

    def read_fasta(self, path: str) -> bytes:
        return read_fasta_bytes(path)

    @_retry_after_oom
    def kmer_set_profile(self, fasta_paths: Sequence[str], kmer_size: int, num_bits: int, modulus: int):
        """Per genome ONE filter of `num_bits` bits over the k-mers whose hash modulo `modulus` is below `num_bits`
        (K1), then their popcounts and pairwise intersections (K3 all pairs) and the number of bits set in at least one /
        at least two of them (msbt_bf_occurrence): (pop[n], inter[n][n], ge1, ge2) as Python integers.  The FASTA text
        stays on the device between calls with the same path list (kopt.py walks k over a range)."""
        key = tuple(fasta_paths)
        if getattr(self, "_set_text_key", None) != key:
            texts = read_files(fasta_paths)
            sizes = [len(t) for t in texts]
            h_text = torch.empty(max(sum(sizes), 1), dtype=torch.uint8, pin_memory=True)
            hv, off = h_text.numpy(), 0
            for t in texts:
                hv[off: off + len(t)] = np.frombuffer(t, dtype=np.uint8)
                off += len(t)
            self._set_text = (h_text.to(self.ctx.device, non_blocking=True), sizes)
            self._set_text_key = key
        d_text, sizes = self._set_text
        batch = self.ctx.pack_device_text(d_text, sizes, kmer_size)
        filters = self.ctx.build_filters(batch, kmer_size, num_bits, min_abund=1, modulus=modulus)
        rows = [filters[i] for i in range(len(sizes))]
        inter = self.ctx.bf_allpairs(rows, num_bits)
        occ = self.ctx.bf_occurrence(rows, num_bits)
        inter_h, occ_h = inter.cpu().tolist(), occ.cpu().tolist()
        return [inter_h[i][i] for i in range(len(sizes))], inter_h, occ_h[0], occ_h[1]

    def kmer_set_profile_done(self) -> None:
        self._set_text_key = self._set_text = None

    # ------------------------------------------------------------------ bfoperate --or
    @_retry_after_oom
    def union(self, src_paths: Sequence[str], out_path: str) -> None:
        out = None
        step = self._chunk()
        for lo in range(0, len(src_paths), step):  # OR-accumulate chunk by chunk
            rows = [self.filter(p) for p in src_paths[lo: lo + step]]
            out = self.ctx.bf_or(rows if out is None else [out] + rows, self.num_bits)
        if self.stride > self.words:
            out[self.words:] = 0
        self._write(out_path, out)
        self._put(out_path, out)

    def union_many(self, jobs: Sequence[Tuple[Sequence[str], str]]) -> None:
        """Many independent `bfoperate --or` calls (all clusters of one taxonomic level, core.py:1408-1427 runs them one
        spawn after the other): the OR kernels are queued back to back, the results leave through a small ring of pinned
        buffers on a side stream and a thread pool writes the `.bf` files while the next unions run.  A job with a single
        source is the byte copy of core.py:3781-3789."""
        if not jobs:
            return
        header = bffile.BloomHeader(kmer_size=self.k, num_bits=self.num_bits)
        ring = [torch.empty(self.words, dtype=torch.int64, pin_memory=True) for _ in range(min(4, len(jobs)))]
        free: "list[torch.Tensor]" = list(ring)
        import threading
        cond = threading.Condition()
        main = torch.cuda.current_stream(self.ctx.device)
        side = torch.cuda.Stream(self.ctx.device)
        step = self._chunk()
        with ThreadPoolExecutor(max_workers=min(4, os.cpu_count() or 1)) as writers:
            pending = []
            for srcs, out_path in jobs:
                out = None
                if len(srcs) == 1:
                    out = self.filter(srcs[0]).clone()
                else:
                    for lo in range(0, len(srcs), step):
                        rows = [self.filter(p) for p in srcs[lo: lo + step]]
                        out = self.ctx.bf_or(rows if out is None else [out] + rows, self.num_bits)
                    if self.stride > self.words:
                        out[self.words:] = 0
                with cond:
                    while not free:
                        cond.wait()
                    host = free.pop()
                built = torch.cuda.Event()
                built.record(main)
                with torch.cuda.stream(side):
                    side.wait_event(built)
                    host.copy_(out[: self.words], non_blocking=True)
                    done = torch.cuda.Event()
                    done.record(side)
                out.record_stream(side)

                def write(out_path=out_path, host=host, done=done):
                    try:
                        done.synchronize()
                        bffile.write_bf(out_path, header, host.numpy().view(np.uint64))
                    finally:
                        with cond:
                            free.append(host)
                            cond.notify()
                pending.append(writers.submit(write))
                self._put(out_path, out)
            for job in pending:
                job.result()

    def copy(self, src_path: str, out_path: str) -> None:
        t = self.filter(src_path).clone()
        self._write(out_path, t)
        self._put(out_path, t)

    # ------------------------------------------------------------------ dumpbf --show:density
    @_retry_after_oom
    def popcounts(self, paths: Sequence[str]) -> List[int]:
        out: List[torch.Tensor] = []
        step = self._chunk()
        for lo in range(0, len(paths), step):
            rows = [self.filter(p) for p in paths[lo: lo + step]]
            out.append(self.ctx.bf_popcount(rows, self.num_bits))
        return [int(x) for x in torch.cat(out).cpu()] if out else []

    # ------------------------------------------------------------------ bfdistance
    @_retry_after_oom
    def dist_one_vs_many(self, focus_path: str, target_paths: Sequence[str]) -> List[Tuple[int, int]]:
        """[(intersection, union)] of the focus against every target, in target order."""
        if not target_paths:
            return []
        focus = self.filter(focus_path)
        inter, union = [], []
        step = self._chunk()
        for lo in range(0, len(target_paths), step):
            i, u = self.ctx.bf_dist(focus, [self.filter(p) for p in target_paths[lo: lo + step]], self.num_bits)
            inter.append(i)
            union.append(u)
        both = torch.stack([torch.cat(inter), torch.cat(union)]).cpu()
        return list(zip(both[0].tolist(), both[1].tolist()))

    @_retry_after_oom
    def dist_pairs(self, pairs: Sequence[Tuple[str, str]]) -> List[Tuple[int, int]]:
        """[(intersection, union)] of an arbitrary list of (path, path) pairs: ONE launch and one device->host copy per
        chunk of pairs (the batch form of Database.dist that profile_many ends with)."""
        if not pairs:
            return []
        inter, union = [], []
        step = max(1, self._chunk() // 2)
        for lo in range(0, len(pairs), step):
            part = pairs[lo: lo + step]
            a = [self.filter(p) for p, _ in part]
            b = [self.filter(q) for _, q in part]
            i, u = self.ctx.bf_dist_pairs(a, b, self.num_bits)
            inter.append(i)
            union.append(u)
        both = torch.stack([torch.cat(inter), torch.cat(union)]).cpu()
        return list(zip(both[0].tolist(), both[1].tolist()))

    def _all_pairs_device(self, paths: Sequence[str]) -> torch.Tensor:
        n = len(paths)
        step = self._chunk(2)
        if n <= step:
            rows = [self.filter(p) for p in paths]
            route = os.environ.get("MSBT_ALLPAIRS_ROUTE", "auto")  # bit | probe | auto
            if route == "auto" and n >= PROBE_MIN_FILTERS and self.num_bits <= (1 << 32):
                free, _total = torch.cuda.mem_get_info(self.ctx.device)
                total_pop = int(self.ctx.bf_popcount(rows, self.num_bits).sum().item())
                route = allpairs_route(n, self.num_bits, total_pop, self.ctx.sliced_row_words(n), free // 2)
            if route == "probe" and self.num_bits <= (1 << 32):
                free, _total = torch.cuda.mem_get_info(self.ctx.device)
                return self.ctx.bf_allpairs_probe(rows, self.num_bits, slice_bytes=max(1 << 28, free // 2))
            return self.ctx.bf_allpairs(rows, self.num_bits)
        # more filters than may be resident at once: tile the matrix; only two half-chunks are alive at a time
        half = max(1, step // 2)
        blocks = [(lo, min(lo + half, n)) for lo in range(0, n, half)]
        out = torch.zeros((n, n), dtype=torch.int64, device=self.ctx.device)
        for bi, (i0, i1) in enumerate(blocks):
            rows_i = [self.filter(paths[t]) for t in range(i0, i1)]
            out[i0:i1, i0:i1] = self.ctx.bf_allpairs(rows_i, self.num_bits)
            for j0, j1 in blocks[bi + 1:]:
                rect = self.ctx.bf_rect(rows_i, [self.filter(paths[t]) for t in range(j0, j1)], self.num_bits)
                out[i0:i1, j0:j1] = rect
                out[j0:j1, i0:i1] = rect.t()
        return out

    @_retry_after_oom
    def dist_all_pairs(self, paths: Sequence[str]) -> np.ndarray:
        """Dense intersection matrix; diagonal = popcounts (union_ij = I_ii + I_jj - I_ij)."""
        return self._all_pairs_device(paths).cpu().numpy()

    @_retry_after_oom
    def dist_all_pairs_many(self, groups: Sequence[Sequence[str]]) -> List[np.ndarray]:
        """dist_all_pairs for many clusters: every all-pairs launch is queued first, the matrices come back in ONE
        device->host copy per cache-sized run of clusters (one synchronisation instead of one per cluster)."""
        out: List[np.ndarray] = []
        run: List[torch.Tensor] = []
        run_sizes: List[int] = []
        alive = 0

        def flush():
            nonlocal run, run_sizes, alive
            if run:
                flat = torch.cat([m.reshape(-1) for m in run]).cpu().numpy()
                off = 0
                for n in run_sizes:
                    out.append(flat[off: off + n * n].reshape(n, n))
                    off += n * n
            run, run_sizes, alive = [], [], 0

        budget = self._chunk(2)
        for g in groups:
            if alive and alive + len(g) > budget:
                flush()
            run.append(self._all_pairs_device(g))
            run_sizes.append(len(g))
            alive += len(g)
        flush()
        return out

    # ------------------------------------------------------------------ query --distinctkmers
    @_retry_after_oom
    def query_positions_many(self, fasta_paths: Sequence[str]) -> QueryBatch:
        """Distinct hash positions of ALL k-mers of every file (no min filter), ascending, on the device.
        They are the set bits of the file's own min=1 filter (App. A6), so per chunk of files: one FASTA ingest, ONE K1
        launch and one batched ordered extraction (three launches)."""
        chunks: List[torch.Tensor] = []
        offsets = [0]
        for lo in range(0, len(fasta_paths), self.sketch_batch):
            texts = read_files(fasta_paths[lo: lo + self.sketch_batch])
            batch = self.ctx.pack_fasta_device(texts, self.k)
            filt = self.ctx.build_filters(batch, self.k, self.num_bits, min_abund=1)
            # a file of n bytes holds fewer than n k-mers: an upper bound that needs no sizing pass
            cap = min(sum(len(t) for t in texts), len(texts) * self.num_bits)
            pos, offs = self.ctx.extract_positions_packed([filt[i] for i in range(len(texts))], self.num_bits, cap=max(cap, 1))
            chunks.append(pos[: offs[-1]])
            base = offsets[-1]
            offsets.extend(base + o for o in offs[1:])
        if not chunks:
            return QueryBatch(torch.zeros(1, dtype=torch.int32, device=self.ctx.device), [0])
        positions = chunks[0] if len(chunks) == 1 else torch.cat(chunks)
        if positions.numel() == 0:
            positions = torch.zeros(1, dtype=torch.int32, device=self.ctx.device)
        return QueryBatch(positions, offsets)

    def query_positions(self, fasta_path: str) -> QueryBatch:
        return self.query_positions_many([fasta_path])

    def query_totals(self, qb: QueryBatch) -> List[int]:
        return [qb.total(i) for i in range(len(qb))]

    @_retry_after_oom
    def query_hits_many(self, qb: QueryBatch, idxs: Sequence[int], leaf_paths: Sequence[str]) -> torch.Tensor:
        """hits[i][l] of queries `idxs` of the batch against the leaves of one flat tree, LEFT ON THE DEVICE (int32
        [len(idxs), len(leaf_paths)]); nothing synchronises -- fetch_hits() brings many of them back in one copy."""
        nq, nl = len(idxs), len(leaf_paths)
        if nq == 0 or nl == 0:
            return torch.zeros((nq, nl), dtype=torch.int32, device=self.ctx.device)
        ranges = qb.ranges(idxs)
        n_pos = sum(e - b for b, e in ranges)
        sliced = self._sliced_tree(leaf_paths, n_pos)
        if sliced is not None:
            self.sliced_probes += nq
            return self.ctx.query_sliced_ranges(sliced, nl, 0, self.num_bits, qb.positions, ranges)
        step = self._chunk()
        if nl <= step:
            return self.ctx.query_ranges([self.filter(p) for p in leaf_paths], qb.positions, ranges)
        parts = [self.ctx.query_ranges([self.filter(p) for p in leaf_paths[lo: lo + step]], qb.positions, ranges)
                 for lo in range(0, nl, step)]
        return torch.cat(parts, dim=1)

    def fetch_hits(self, pending: Sequence[torch.Tensor]) -> List[np.ndarray]:
        """Device hit matrices -> host, in ONE device->host copy (one synchronisation per tree level of profile_many)."""
        if not pending:
            return []
        flat = torch.cat([t.reshape(-1) for t in pending]).cpu().numpy()
        out, off = [], 0
        for t in pending:
            n = t.numel()
            out.append(flat[off: off + n].reshape(t.shape))
            off += n
        return out

    def query_hits(self, positions: QueryBatch, leaf_paths: Sequence[str]) -> Tuple[List[int], int]:
        """(hits per leaf, total) for one query against the leaves of a flat tree."""
        total = positions.total(0)
        if not leaf_paths:
            return [], total
        hits = self.fetch_hits([self.query_hits_many(positions, [0], leaf_paths)])[0]
        return [int(x) for x in hits[0]], total

    def _sliced_tree(self, leaf_paths: Sequence[str], n_positions: int = 0) -> Optional[torch.Tensor]:
        """Bit-sliced matrix of a flat tree's leaves, or None when the row-major probe should serve the query.
        The row-major probe reads one 32-byte sector per (position, leaf); the bit-sliced one reads 1/8 byte per
        (position, leaf) but needs the leaves transposed first (one read of every leaf + one write of the matrix).
        A tree with many leaves gets transposed as soon as that pays: when the row-major probe of THIS call would move
        more bytes than the transposition plus the sliced probe (a batch of MAGs from profile_many), or on the tree's
        second query (`profile` walks the same trees for every input genome).  The matrix is kept (LRU, bounded in
        bytes) and dropped when one of its filters is rewritten."""
        n = len(leaf_paths)
        if n < self.sliced_min_leaves:
            return None
        key = tuple(os.path.abspath(p) for p in leaf_paths)
        t = self._sliced.get(key)
        if t is not None:
            self._sliced.move_to_end(key)
            return t
        seen = self._tree_queries.pop(key, 0) + 1
        self._tree_queries[key] = seen
        while len(self._tree_queries) > 4096:
            self._tree_queries.popitem(last=False)
        row_words = self.ctx.sliced_row_words(n)
        nbytes = self.num_bits * row_words * 4
        if nbytes > self.sliced_cache_bytes or n > self._chunk():
            return None
        rowmajor = n_positions * n * 32
        sliced_cost = 2 * n * self.row_bytes + n_positions * row_words * 4
        if seen < 2 and rowmajor <= sliced_cost:
            return None
        t = self.ctx.bitslice([self.filter(p) for p in key], 0, self.num_bits)
        self._sliced[key] = t
        self._sliced_bytes += nbytes
        while self._sliced_bytes > self.sliced_cache_bytes and len(self._sliced) > 1:
            _, old = self._sliced.popitem(last=False)
            self._sliced_bytes -= old.numel() * 4
        return t
