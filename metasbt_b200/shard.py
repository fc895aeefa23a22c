"""Multi-GPU partitioning of the hot path: one process per GPU, torch.distributed for the
plumbing.  Pure host logic + collectives on whatever backend the process group uses (NCCL on the box,
gloo in the CPU tests), so the partition / merge arithmetic is testable without a GPU.

* K1 construction  : genomes are independent -> contiguous blocks balanced by k-mer count, NO collective.
* K3 / K3' counts  : filters sharded by bit range -> partial counts -> all_reduce(SUM) of u64 counters.
* K4 query         : bit-sliced leaf matrix sharded by contiguous row (bit) range; every rank extracts and probes
                     the positions that fall into its rows -> partial hits[B][L] -> all_reduce(SUM)
                     (query_sharded).
* K2 union         : OR is not an NCCL reduction op.  Partials are merged by bit range: all_to_all of
                     range r of every partial to rank r, local OR, all_gather of the merged ranges.
"""

from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def genome_blocks(n_bases: Sequence[int], world: int, k: int = 1) -> List[Tuple[int, int]]:
    """Contiguous [begin, end) genome ranges per rank, balanced by the number of k-mers (greedy prefix cut)."""
    work = [max(int(n) - k + 1, 0) + 1 for n in n_bases]
    total = sum(work)
    out, begin, acc = [], 0, 0
    for r in range(world):
        target = total * (r + 1) / world
        end = begin
        while end < len(work) and (acc + work[end] <= target or (r == world - 1)):
            acc += work[end]
            end += 1
        # never leave later ranks without genomes they could have had
        out.append((begin, end))
        begin = end
    if out:
        out[-1] = (out[-1][0], len(work))
    return out


def bit_ranges(num_bits: int, world: int, align: int = 128) -> List[Tuple[int, int]]:
    """Contiguous [begin, end) bit ranges per rank; boundaries are multiples of `align` bits (128 = one
    uint4) so that every shard of a filter row stays 16-byte aligned."""
    units = (num_bits + align - 1) // align
    out = []
    for r in range(world):
        b = units * r // world * align
        e = min(units * (r + 1) // world * align, num_bits) if r < world - 1 else num_bits
        out.append((min(b, num_bits), e))
    return out


def all_reduce_counts(partial: torch.Tensor, group=None) -> torch.Tensor:
    """Sum partial popcounts / hit counts of bit-range shards across ranks (in place)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(partial, op=dist.ReduceOp.SUM, group=group)
    return partial


def query_sharded(ctx, sliced: torch.Tensor, n_leaves: int, num_bits: int, query_filters: Sequence[torch.Tensor],
                  cap: int | None = None, group=None, rank: int | None = None, world: int | None = None) -> torch.Tensor:
    """K4 across ranks.  `sliced` holds this rank's rows of the bit-sliced leaf matrix (ops.Context.bitslice on
    bit_ranges(num_bits, world)[rank]); `query_filters` are the queries' own min=1 filters (whole, on every rank).
    Each rank extracts only the positions of ITS bit range (word sub-views of the query filters: the ranges are 128-bit
    aligned), probes its shard with them and the partial hit matrices are summed -> int32[q, l] on every rank.
    With explicit `rank` / `world` the partial matrix of that shard is returned and no collective runs (tests)."""
    partial_only = rank is not None and world is not None
    if not partial_only:
        world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        rank = dist.get_rank(group) if world > 1 else 0
    rb, re_ = bit_ranges(num_bits, world)[rank]
    w0, w1 = rb // 64, (re_ + 63) // 64
    pos, offs = ctx.extract_positions_packed([f[w0:w1] for f in query_filters], (w1 - w0) * 64, cap)
    hits = ctx.query_sliced_packed(sliced, n_leaves, 0, re_ - rb, pos, offs)  # positions are relative to rb
    return hits if partial_only else all_reduce_counts(hits, group)


def query_blocks(n_queries: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous [begin, end) blocks of a query batch per rank (the rank that HASHES those MAGs)."""
    return [(n_queries * r // world, n_queries * (r + 1) // world) for r in range(world)]


def exchange_plan(counts_all: Sequence[Sequence[Sequence[int]]], rank: int) -> Tuple[List[int], List[int], List[Tuple[int, int]]]:
    """Host-side plan of the position exchange.  counts_all[s][q][d] = number of positions of the q-th MAG hashed by rank s
    that fall into rank d's bit range.  Returns for `rank`: (send split sizes per destination, receive split sizes per
    source, and the [begin, end) range inside the receive buffer of every query of the batch in global batch order)."""
    world = len(counts_all)
    send = [sum(int(c[d]) for c in counts_all[rank]) for d in range(world)]
    recv = [sum(int(c[rank]) for c in counts_all[s]) for s in range(world)]
    ranges, cursor = [], 0
    for s in range(world):
        for c in counts_all[s]:
            ranges.append((cursor, cursor + int(c[rank])))
            cursor += int(c[rank])
    return send, recv, ranges


class QueryExchange:
    """What query_exchange_prepare hands to query_exchange_finish: the positions of the WHOLE batch inside this rank's bit
    range (device, relative to the range start) and each query's [begin, end) inside them."""

    def __init__(self, recv: torch.Tensor, q_ranges: List[Tuple[int, int]], n_received: int):
        self.recv, self.q_ranges, self.n_received = recv, q_ranges, n_received


def query_exchange_prepare(ctx, num_bits: int, local_filters: Sequence[torch.Tensor], group=None, cap: int | None = None,
                           batch_size: int | None = None, events: list | None = None) -> QueryExchange:
    """Steps 1-3 of query_exchange_sharded (extraction, counts, all_to_all) on torch's CURRENT stream -- callers that
    pipeline batches run it on a side stream while the previous batch is being probed."""
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank(group) if world > 1 else 0
    dev = ctx.device
    n_local = len(local_filters)
    ranges_bits = bit_ranges(num_bits, world)
    words = [(b // 64, (e + 63) // 64) for b, e in ranges_bits]
    if events:
        events[0].record()
    # 1. positions, destination-major, straight into the send buffer
    equal = len({w1 - w0 for w0, w1 in words}) == 1
    if n_local == 0:
        counts_h: List[List[int]] = []
        send = torch.empty(1, dtype=torch.int32, device=dev)
    else:
        if cap is None:
            cap = int(ctx.bf_popcount(list(local_filters), num_bits).sum().item())
        send = torch.empty(max(cap, 1), dtype=torch.int32, device=dev)
        if equal:
            views = [f[w0:w1] for w0, w1 in words for f in local_filters]
            offs = ctx.extract_positions_into(views, 64 * (words[0][1] - words[0][0]), send).cpu().tolist()
        else:  # ragged ranges: one extraction per destination
            offs, cursor = [0], 0
            for w0, w1 in words:
                o = ctx.extract_positions_into([f[w0:w1] for f in local_filters], 64 * (w1 - w0), send[cursor:]).cpu().tolist()
                offs.extend(cursor + x for x in o[1:])
                cursor += o[-1]
        if offs[-1] > send.numel():
            raise ValueError("query_exchange: `cap` is smaller than the number of positions of the local queries")
        counts_h = [[offs[d * n_local + q + 1] - offs[d * n_local + q] for d in range(world)] for q in range(n_local)]
    # 2. everybody's counts
    if world == 1:
        gathered: List[List[List[int]]] = [counts_h]
    elif batch_size is not None:
        blocks = query_blocks(batch_size, world)
        n_max = max(e - b for b, e in blocks)
        mine = torch.zeros((n_max, world), dtype=torch.int64)
        if n_local:
            mine[:n_local] = torch.tensor(counts_h, dtype=torch.int64)
        mine = mine.to(dev)
        allc = torch.empty((world, n_max, world), dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(allc, mine, group=group)
        allc_h = allc.cpu().tolist()
        gathered = [allc_h[r][: blocks[r][1] - blocks[r][0]] for r in range(world)]
    else:
        gathered = [None] * world  # type: ignore[list-item]
        dist.all_gather_object(gathered, counts_h, group=group)
    send_sizes, recv_sizes, q_ranges = exchange_plan(gathered, rank)
    if events:
        events[1].record()
    # 3. exchange
    if world > 1:
        recv = torch.empty(max(sum(recv_sizes), 1), dtype=torch.int32, device=dev)
        dist.all_to_all_single(recv[: sum(recv_sizes)], send[: sum(send_sizes)], recv_sizes, send_sizes, group=group)
    else:
        recv = send
    if events:
        events[2].record()
    return QueryExchange(recv, q_ranges, sum(recv_sizes))


def query_exchange_finish(ctx, sliced: torch.Tensor, n_leaves: int, num_bits: int, ex: QueryExchange, group=None,
                          events: list | None = None, reduce: bool = True) -> torch.Tensor:
    """Steps 4-5: probe this rank's shard with the exchanged positions, sum the partial hit matrices."""
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank(group) if world > 1 else 0
    rb, re_ = bit_ranges(num_bits, world)[rank]
    ex.recv.record_stream(torch.cuda.current_stream(ctx.device)) if ex.recv.is_cuda else None
    hits = ctx.query_sliced_ranges(sliced, n_leaves, 0, re_ - rb, ex.recv, ex.q_ranges)  # row indices relative to the range
    if events:
        events[3].record()
    if reduce:
        all_reduce_counts(hits, group)
    if events:
        events[4].record()
    return hits


def query_exchange_sharded(ctx, sliced: torch.Tensor, n_leaves: int, num_bits: int, local_filters: Sequence[torch.Tensor],
                           group=None, timings: dict | None = None, cap: int | None = None,
                           batch_size: int | None = None) -> torch.Tensor:
    """K4 across ranks with the HASHING sharded as well (second option).  Every rank has hashed only
    its block of the MAG batch (query_blocks): `local_filters` are those MAGs' own min=1 filters.  Per step:
      1. ONE batched ordered extraction over the word sub-views (destination-major) of the local filters: the positions of
         every local MAG inside every rank's bit range, relative to the range start, land in the send buffer in exactly the
         order the exchange wants; the extraction's offsets give the split sizes (one small device->host copy);
      2. all_gather of the per-(MAG, destination) counts (a fixed-size tensor when `batch_size` is given);
      3. all_to_all of the position lists (rank r receives, MAG by MAG, the positions inside ITS rows);
      4. one bit-sliced probe of the local shard with the positions of the WHOLE batch -> partial hits[B][L];
      5. all_reduce(SUM) of the partial hit matrices.
    `cap`: upper bound on the number of positions of the local MAGs (e.g. their summed lengths); None sizes the send
    buffer with a popcount pass.  Returns int32[B, L] on every rank, B = the batch size over all ranks, in rank order.
    query_exchange_prepare (1-3) and query_exchange_finish (4-5) are the two halves, for callers that overlap the
    preparation of the next batch with the probe of the current one."""
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)] if timings is not None else None
    ex = query_exchange_prepare(ctx, num_bits, local_filters, group, cap, batch_size, ev)
    hits = query_exchange_finish(ctx, sliced, n_leaves, num_bits, ex, group, ev)
    if ev:
        torch.cuda.synchronize(ctx.device)
        for name, i in (("ms_extract_and_counts", 0), ("ms_all_to_all", 1), ("ms_probe", 2), ("ms_all_reduce", 3)):
            timings[name] = timings.get(name, 0.0) + ev[i].elapsed_time(ev[i + 1])
        timings["positions_in_shard"] = timings.get("positions_in_shard", 0) + ex.n_received
    return hits


def ring_schedule(rank: int, world: int) -> List[Tuple[int, int, str]]:
    """Which (peer to receive from, peer to send to, role) a rank meets in the half ring of all_pairs_ring.  Every unordered
    pair of DIFFERENT blocks is computed exactly once: at step s = 1 .. floor(world / 2) rank r receives block (r + s) and
    sends its own to (r - s); when world is even, the last step pairs r with r + world/2 in both directions, so the lower
    rank takes the first half of its rows against the whole peer block ("low") and the upper one all of its rows against
    the second half of the peer's ("high").  role is "full", "low" or "high"."""
    out = []
    for s in range(1, world // 2 + 1):
        src, dst = (rank + s) % world, (rank - s) % world
        role = "full"
        if world % 2 == 0 and s == world // 2:
            role = "low" if rank < src else "high"
        out.append((src, dst, role))
    return out


def all_pairs_ring(ctx, mat: torch.Tensor, num_bits: int, group=None, chunk: int = 128) -> torch.Tensor:
    """All-pairs intersections of a filter set that is spread over the ranks (each holds a block of rows `mat`
    [n_local, stride] int64) and does not fit one GPU -- the `dist` loop of Entry.get_boundaries (core.py:4036-4081) over
    e.g. 10,000 genomes.  Local block: one square all-pairs launch.  Other blocks: half ring (ring_schedule): the peer's
    filters arrive `chunk` at a time over NCCL send/recv and meet the local rows in one rectangular launch each
    (msbt_bf_and_popcount_rect).  Every pair is computed once.  Returns int64 [n_local, n_total] with -1 where the pair
    belongs to the other side; all_pairs_assemble() gives the full symmetric matrix."""
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank(group) if world > 1 else 0
    n_local, stride = int(mat.shape[0]), int(mat.shape[1])
    counts = [n_local]
    if world > 1:
        t = torch.tensor([n_local], dtype=torch.int64, device=mat.device)
        allc = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allc, t, group=group)
        counts = [int(c.item()) for c in allc]
    offs = [0]
    for c in counts:
        offs.append(offs[-1] + c)
    out = torch.full((n_local, offs[-1]), -1, dtype=torch.int64, device=mat.device)
    rows = [mat[i] for i in range(n_local)]
    if n_local:
        out[:, offs[rank]: offs[rank + 1]] = ctx.bf_allpairs(rows, num_bits)
    for src, dst, role in ring_schedule(rank, world):
        n_src, n_dst = counts[src], counts[dst]
        # what travels: the whole block, except towards the "low" side of an opposite pair, which only needs ... the whole
        # block too (it takes all of the peer's filters); the "high" side needs the second half of the lower rank's rows
        send_lo, send_hi = (0, n_local)
        recv_lo, recv_hi = (0, n_src)
        if role == "low":      # I send my second half (the peer is "high"), I receive the peer's whole block
            send_lo = n_local // 2
        elif role == "high":   # I send my whole block (the peer is "low"), I receive the peer's second half
            recv_lo = n_src // 2
        my_rows = rows[: n_local // 2] if role == "low" else rows
        row_hi = n_local // 2 if role == "low" else n_local
        n_send, n_recv = send_hi - send_lo, recv_hi - recv_lo
        steps = max((n_send + chunk - 1) // chunk, (n_recv + chunk - 1) // chunk)
        for c in range(steps):
            sb, se = min(send_lo + c * chunk, send_hi), min(send_lo + (c + 1) * chunk, send_hi)
            rb, re_ = min(recv_lo + c * chunk, recv_hi), min(recv_lo + (c + 1) * chunk, recv_hi)
            buf = torch.empty((re_ - rb, stride), dtype=mat.dtype, device=mat.device)
            reqs = []
            if se > sb:
                reqs.append(dist.P2POp(dist.isend, mat[sb:se].contiguous(), dst if group is None else dist.get_global_rank(group, dst), group))
            if re_ > rb:
                reqs.append(dist.P2POp(dist.irecv, buf, src if group is None else dist.get_global_rank(group, src), group))
            for q in (dist.batch_isend_irecv(reqs) if reqs else []):
                q.wait()
            if re_ > rb and my_rows:
                block = ctx.bf_rect(my_rows, [buf[i] for i in range(re_ - rb)], num_bits)
                out[:row_hi, offs[src] + rb: offs[src] + re_] = block
    return out


def all_pairs_probe(ctx, mat: torch.Tensor, num_bits: int, group=None, chunk: int = 256, batch: int = 256,
                    slice_bytes: Optional[int] = None) -> torch.Tensor:
    """All-pairs intersections of MANY SPARSE filters spread over the ranks (each holds a block of rows `mat`
    [n_local, stride] int64), the K4 way (ops.bf_allpairs_probe): |A and B| over the whole filter is the sum over bit ranges
    of |A and B| within the range, so rank r receives range r (bit_ranges) of EVERY filter -- `chunk` filters per peer and
    round over send/recv --, transposes those sub-filters into its shard of the bit-sliced matrix, lets every sub-filter's
    set bits probe it, and the partial [n_total, n_total] matrices are summed with one all_reduce.  Returns the full
    symmetric int64 matrix (diagonal = popcounts) on every rank.  Work per rank ~ (set bits of all filters) / world rows of
    n_total bits, against n_total^2 / (2 world) whole-filter passes for all_pairs_ring."""
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank(group) if world > 1 else 0
    n_local = int(mat.shape[0])
    words = (num_bits + 63) // 64
    kw = dict() if slice_bytes is None else dict(slice_bytes=slice_bytes)
    if world == 1:
        return ctx.bf_allpairs_probe([mat[i] for i in range(n_local)], num_bits, batch=batch, **kw)
    t = torch.tensor([n_local], dtype=torch.int64, device=mat.device)
    allc = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(allc, t, group=group)
    counts = [int(c.item()) for c in allc]
    offs = [0]
    for c in counts:
        offs.append(offs[-1] + c)
    ranges = bit_ranges(num_bits, world)                      # multiples of 128 bits: word-aligned, 16-byte aligned
    wr = [(b // 64, min(words, (e + 63) // 64)) for b, e in ranges]
    my_b, my_e = ranges[rank]
    my_w = wr[rank][1] - wr[rank][0]
    stride_r = max(2, (my_w + 1) // 2 * 2)                    # rows of the sub-filter matrix stay 16-byte aligned
    sub = torch.zeros((offs[-1], stride_r), dtype=mat.dtype, device=mat.device)
    glob = (lambda r: r) if group is None else (lambda r: dist.get_global_rank(group, r))
    for c0 in range(0, max(counts), chunk):
        reqs, landed = [], []
        for p in range(world):
            sb, se = min(c0, n_local), min(c0 + chunk, n_local)           # my rows of this round
            rb, re_ = min(c0, counts[p]), min(c0 + chunk, counts[p])      # peer p's rows of this round
            if p == rank:
                if se > sb and my_w:
                    sub[offs[rank] + sb: offs[rank] + se, :my_w] = mat[sb:se, wr[rank][0]: wr[rank][1]]
                continue
            pw = wr[p][1] - wr[p][0]
            if se > sb and pw:
                reqs.append(dist.P2POp(dist.isend, mat[sb:se, wr[p][0]: wr[p][1]].contiguous(), glob(p), group))
            if re_ > rb and my_w:
                buf = torch.empty((re_ - rb, my_w), dtype=mat.dtype, device=mat.device)
                reqs.append(dist.P2POp(dist.irecv, buf, glob(p), group))
                landed.append((offs[p] + rb, offs[p] + re_, buf))
        for q in (dist.batch_isend_irecv(reqs) if reqs else []):
            q.wait()
        for lo, hi, buf in landed:
            sub[lo:hi, :my_w] = buf
    partial = ctx.bf_allpairs_probe([sub[i] for i in range(offs[-1])], my_e - my_b, batch=batch, **kw) if my_e > my_b else \
        torch.zeros((offs[-1], offs[-1]), dtype=torch.int64, device=mat.device)
    dist.all_reduce(partial, op=dist.ReduceOp.SUM, group=group)
    return partial


def all_pairs_assemble(local: torch.Tensor, group=None) -> torch.Tensor:
    """Full symmetric [n_total, n_total] matrix on every rank from the per-rank row blocks of all_pairs_ring (small n:
    tests, reports)."""
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    if world == 1:
        return local
    n_total = int(local.shape[1])
    t = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
    allc = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(allc, t, group=group)
    counts = [int(c.item()) for c in allc]
    pad = max(counts)
    mine = torch.full((pad, n_total), -1, dtype=torch.int64, device=local.device)
    mine[: local.shape[0]] = local
    blocks = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(blocks, mine, group=group)
    full = torch.cat([b[:c] for b, c in zip(blocks, counts)])
    return torch.where(full >= 0, full, full.t())


def or_merge(partial: torch.Tensor, num_bits: int, group=None, ctx=None) -> torch.Tensor:
    """Bitwise-OR merge of one partial filter per rank (int64 words, same length everywhere).
    all_to_all by bit range, local k-way OR of the received ranges (msbt_bf_or_reduce through `ctx` on the GPU; the
    gloo CPU tests, which have no device, fall back to torch.bitwise_or for this one elementwise step), then all_gather
    of the merged ranges.  Returns the full merged filter on every rank."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return partial
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    words = partial.numel()
    ranges = [(b // 64, (e + 63) // 64) for b, e in bit_ranges(words * 64, world)]
    send = [partial[b:e].contiguous() for b, e in ranges]
    mine = ranges[rank][1] - ranges[rank][0]
    recv = [torch.empty(mine, dtype=partial.dtype, device=partial.device) for _ in range(world)]
    dist.all_to_all(recv, send, group=group) if partial.is_cuda else _all_to_all_gloo(recv, send, group)
    if partial.is_cuda:
        if ctx is None:
            raise ValueError("or_merge on CUDA tensors needs the ops.Context that owns the device (msbt_bf_or_reduce)")
        merged = ctx.bf_or(recv, 64 * mine)[:mine]  # fresh allocations: 16-byte aligned rows, as msbt_bf_or_reduce wants
    else:
        merged = recv[0]
        for r in range(1, world):
            merged = torch.bitwise_or(merged, recv[r])
    pieces = [torch.empty(e - b, dtype=partial.dtype, device=partial.device) for b, e in ranges]
    dist.all_gather(pieces, merged, group=group) if len({e - b for b, e in ranges}) == 1 else _all_gather_uneven(pieces, merged, group)
    return torch.cat(pieces)


def _all_to_all_gloo(recv, send, group):
    """gloo has no all_to_all: emulate with world broadcasts-free point-to-point rounds."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    recv[rank].copy_(send[rank])
    for shift in range(1, world):
        dst, src = (rank + shift) % world, (rank - shift) % world
        reqs = [dist.isend(send[dst], dst, group=group), dist.irecv(recv[src], src, group=group)]
        for q in reqs:
            q.wait()


def _all_gather_uneven(pieces, mine, group):
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    for r in range(world):
        if r == rank:
            pieces[r].copy_(mine)
        dist.broadcast(pieces[r], src=dist.get_global_rank(group, r) if group is not None else r, group=group)
