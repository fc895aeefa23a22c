"""Host FASTA bytes -> Bloom filters in host memory, as a software pipeline over one GPU.

This is the makebf stage end to end (`howdesbt makebf <fasta> --out=<bf>`, core.py:3920-3931) for a stream of
genomes: the reference spawns one process per genome, each reading a FASTA file and writing a `.bf` file.  Here the
genomes move through the device in chunks, three stages overlapping on three CUDA streams:

    stream_in   FASTA text, pinned host -> device                        (5 MB per 5 Mbp genome)
    compute     msbt_fasta_pack_device + msbt_kmer_bloom_build [+ msbt_bf_extract_positions_batch]
    stream_out  the chunk's filters, device -> pinned host slot           (dense: num_bits/8 bytes per genome;
                                                                           sparse: 4 bytes per set bit)
    sink        worker threads consume the slot (write `.bf` files) and hand it back

PCIe is full duplex, so the output copy of chunk c runs under the input copy, ingest and K1 of chunk c + 1; with
`slots` >= 2 the consumer of chunk c runs under all of that.  In `sparse` mode a filter leaves the device as the ascending
list of its set bits (a 5 Mbp genome in a 2^30-bit filter: 20 MB instead of 128 MiB) and the dense words are rebuilt in
the host slot by msbt_positions_expand on the sink's threads; what the sink sees is identical in every mode.  The dense
transport is bound by the copy engine and leaves the host cores idle, the sparse one is bound by the host cores and
leaves the copy engine idle: `hybrid` sends part of every chunk either way.
Pinned buffers and device slots are allocated once per pipeline and reused.
"""

from __future__ import annotations

import threading
import time
from concurrent.futures import ThreadPoolExecutor
from typing import Callable, List, Optional, Sequence

import numpy as np
import torch

from . import ops


class _Slot:
    def __init__(self, ctx: ops.Context, chunk: int, num_bits: int, text_cap: int, sparse: bool):
        self.words = ops.filter_words(num_bits)
        self.d_text = torch.empty(max(text_cap, 1), dtype=torch.uint8, device=ctx.device)
        self.d_filters = ctx.alloc_filters(chunk, num_bits)
        self.h_filters = torch.empty((chunk, self.words), dtype=torch.int64, pin_memory=True)
        self.h_positions: Optional[torch.Tensor] = None
        if sparse:
            # a file of n bytes holds fewer than n k-mers, hence fewer than n set bits
            self.h_positions = torch.empty(max(text_cap, 1), dtype=torch.int32, pin_memory=True)
        self.free = threading.Event()
        self.free.set()
        self.copied = torch.cuda.Event()
        self.positions_copied = torch.cuda.Event()
        self.error: Optional[BaseException] = None


class SketchPipeline:
    """makebf for a stream of genomes whose FASTA text sits in (pinned) host memory; filters are delivered to `sink` in
    host memory, chunk by chunk.  mode: "dense" (filters cross PCIe as they are), "sparse" (position lists cross, dense
    words rebuilt on the host), "hybrid" (both at once: part of every chunk dense on the copy engine, the rest sparse on
    the host cores, the split following their measured costs) or "auto" (sparse when the position lists are expected to
    be under half the dense bytes, else dense)."""

    def __init__(self, ctx: ops.Context, kmer_size: int, num_bits: int, min_abund: int = 1, chunk_genomes: int = 32,
                 slots: int = 2, mode: str = "auto", variant: int = 0, expand_threads: Optional[int] = None):
        if mode not in ("dense", "sparse", "hybrid", "auto"):
            raise ValueError("mode must be dense, sparse, hybrid or auto")
        self.ctx, self.k, self.num_bits, self.min_abund = ctx, int(kmer_size), int(num_bits), int(min_abund)
        self.chunk, self.n_slots, self.mode, self.variant = int(chunk_genomes), int(slots), mode, variant
        self.words = ops.filter_words(self.num_bits)
        self.stream_in = torch.cuda.Stream(ctx.device)
        self.stream_out = torch.cuda.Stream(ctx.device)
        self._slots: List[_Slot] = []
        self._text_cap = 0
        self._sparse = False                     # any position-list transport (sparse or hybrid)
        self._hybrid = mode == "hybrid"          # (auto picks plain sparse: on the boxes measured, the two transports share the
                                                 # host's memory system and hybrid is within +-25 % of sparse, not reliably above)
        self._dense_fraction = 0.3               # hybrid: share of a chunk that travels dense; follows the measured costs
        import os
        # host threads that rebuild dense words (sparse transport): the box's cores are shared by the ranks on it
        ranks_here = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", os.environ.get("WORLD_SIZE", "1")) or 1))
        self._threads = expand_threads or max(2, min(self.chunk, (os.cpu_count() or 2) // ranks_here))
        self._pool = ThreadPoolExecutor(max_workers=self._threads)
        self._sinks = ThreadPoolExecutor(max_workers=self.n_slots)
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    def close(self) -> None:
        self._pool.shutdown(wait=True)
        self._sinks.shutdown(wait=True)
        self._slots = []

    def _prepare(self, offsets: Sequence[int]) -> List[range]:
        n = len(offsets) - 1
        chunks = [range(lo, min(lo + self.chunk, n)) for lo in range(0, n, self.chunk)]
        text_cap = max((offsets[c[-1] + 1] - offsets[c[0]] for c in chunks), default=1)
        total_text = offsets[-1] - offsets[0]
        sparse = self.mode in ("sparse", "hybrid") or (self.mode == "auto" and self.min_abund <= 1
                                                       and total_text * 4 < n * (self.num_bits // 8) // 2)
        if sparse and self.num_bits > (1 << 32):
            sparse = False
        if not self._slots or text_cap > self._text_cap or sparse != self._sparse:
            self._slots = [_Slot(self.ctx, self.chunk, self.num_bits, text_cap, sparse) for _ in range(self.n_slots)]
            self._text_cap, self._sparse = text_cap, sparse
        return chunks

    def run(self, text: torch.Tensor, offsets: Sequence[int], sink: Optional[Callable[[int, np.ndarray], None]] = None,
            on_device: Optional[Callable[[int, torch.Tensor], None]] = None) -> None:
        """`text` (uint8, ideally pinned) holds the FASTA files back to back, genome g is text[offsets[g]:offsets[g+1]].
        sink(first_genome_index, words) is called on a worker thread per chunk with a uint64 [genomes, filter_words] view
        of the slot's host buffer, valid until it returns.  on_device(first_genome_index, filters) is called on the
        calling thread right after the chunk's K1 launch with the slot's device matrix (valid for stream-ordered use
        only: the slot is rebuilt `slots` chunks later).  Returns when every chunk has been consumed."""
        offsets = [int(o) for o in offsets]
        if len(offsets) < 2:
            return
        chunks = self._prepare(offsets)
        main = torch.cuda.current_stream(self.ctx.device)
        jobs = []
        for c, members in enumerate(chunks):
            slot = self._slots[c % self.n_slots]
            slot.free.wait()
            if slot.error is not None:
                err, slot.error = slot.error, None
                raise err
            slot.free.clear()
            try:
                jobs.append(self._submit_chunk(slot, members, text, offsets, sink, on_device, main))
            except BaseException:
                slot.free.set()  # a failed launch must not leave the slot taken: the pipeline stays usable
                raise
        for j in jobs:
            j.result()

    def _submit_chunk(self, slot: "_Slot", members: range, text: torch.Tensor, offsets: List[int], sink, on_device, main):
        ctx = self.ctx
        b, e = offsets[members[0]], offsets[members[-1] + 1]
        sizes = [offsets[g + 1] - offsets[g] for g in members]
        nc = len(members)
        with torch.cuda.stream(self.stream_in):
            slot.d_text[: e - b].copy_(text[b:e], non_blocking=True)
            arrived = torch.cuda.Event()
            arrived.record(self.stream_in)
        self.h2d_bytes += e - b
        main.wait_event(arrived)
        main.wait_event(slot.copied)  # the slot's previous output copy has read d_filters
        batch = ctx.pack_device_text(slot.d_text, sizes, self.k)
        ctx.build_filters(batch, self.k, self.num_bits, min_abund=self.min_abund, out=slot.d_filters, variant=self.variant)
        if on_device is not None:
            on_device(members[0], slot.d_filters[:nc])
        # Output: the first n_dense filters of the chunk cross PCIe as they are (copy engine, no host work), the rest as
        # position lists that host threads expand -- the two transports run side by side (hybrid), the split follows
        # their measured per-genome costs.
        n_dense = nc if not self._sparse else (min(nc, int(round(nc * self._dense_fraction))) if self._hybrid else 0)
        offs = None
        if n_dense < nc:
            sparse_rows = [slot.d_filters[i] for i in range(n_dense, nc)]
            cap = max(sum(sizes[n_dense:]), 1)
            pos, offs = ctx.extract_positions_packed(sparse_rows, self.num_bits, cap=cap)
        built = torch.cuda.Event()
        built.record(main)
        t_dense = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) if 0 < n_dense < nc else None
        with torch.cuda.stream(self.stream_out):
            self.stream_out.wait_event(built)
            if offs is not None:
                slot.h_positions[: offs[-1]].copy_(pos[: offs[-1]], non_blocking=True)
                slot.positions_copied.record(self.stream_out)
                self.d2h_bytes += offs[-1] * 4
            if n_dense:
                if t_dense:
                    t_dense[0].record(self.stream_out)
                slot.h_filters[:n_dense].copy_(slot.d_filters[:n_dense, : self.words], non_blocking=True)
                if t_dense:
                    t_dense[1].record(self.stream_out)
                self.d2h_bytes += n_dense * self.words * 8
            slot.copied.record(self.stream_out)
        if offs is not None:
            pos.record_stream(self.stream_out)
        return self._sinks.submit(self._consume, slot, members[0], nc, n_dense, offs, sink, t_dense)

    def _consume(self, slot: _Slot, first: int, nc: int, n_dense: int, offs: Optional[List[int]], sink, t_dense) -> None:
        try:
            host = slot.h_filters.numpy().view(np.uint64)
            t_expand = 0.0
            if offs is not None:
                slot.positions_copied.synchronize()   # the position lists are here: expand while the dense rows still travel
                hp = slot.h_positions.numpy().view(np.uint32)
                t0 = time.perf_counter()
                list(self._pool.map(lambda i: ops.positions_expand(hp[offs[i]: offs[i + 1]], self.num_bits, host[n_dense + i]),
                                    range(nc - n_dense)))
                t_expand = time.perf_counter() - t0
            slot.copied.synchronize()
            if t_dense is not None and t_expand > 0:
                # per-genome cost of either transport -> the split that would have made both finish together
                c_dense = t_dense[0].elapsed_time(t_dense[1]) / 1e3 / n_dense
                c_sparse = t_expand / (nc - n_dense)
                target = c_sparse / (c_sparse + c_dense)
                self._dense_fraction = min(0.9, max(0.05, 0.5 * self._dense_fraction + 0.5 * target))
            if sink is not None:
                sink(first, host[:nc])
        except BaseException as exc:  # surfaced by run() at the slot's next use or at the end
            slot.error = exc
            raise
        finally:
            slot.free.set()
