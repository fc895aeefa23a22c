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

from typing import List, Sequence, Tuple

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


def or_merge(partial: torch.Tensor, num_bits: int, group=None) -> torch.Tensor:
    """Bitwise-OR merge of one partial filter per rank (int64 words, same length everywhere).
    all_to_all by bit range, local OR (torch.bitwise_or on the received [world, range] block -- a trivially
    memory-bound elementwise op on data already resident; the k-way OR of whole filters is
    msbt_bf_or_reduce), then all_gather of the merged ranges.  Returns the full merged filter on every rank."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return partial
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    words = partial.numel()
    ranges = [(b // 64, (e + 63) // 64) for b, e in bit_ranges(words * 64, world)]
    send = [partial[b:e].contiguous() for b, e in ranges]
    mine = ranges[rank][1] - ranges[rank][0]
    recv = [torch.empty(mine, dtype=partial.dtype, device=partial.device) for _ in range(world)]
    dist.all_to_all(recv, send, group=group) if partial.is_cuda else _all_to_all_gloo(recv, send, group)
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
