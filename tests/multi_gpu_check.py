#!/usr/bin/env python
"""Multi-GPU parity check over NCCL, run under torchrun on N GPUs of one box:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29512 tests/multi_gpu_check.py

* K1: genomes sharded in contiguous blocks, no collective; every rank's filters equal the oracle's.
* K2: per-rank partial unions merged by bit range (all_to_all + local OR + all_gather) == OR of everything.
* K3: AND-popcounts over bit-range shards, all_reduce(SUM) == whole-filter counts.
* K4: bit-sliced leaf matrix sharded by row range, partial hits all_reduce(SUM) == unsharded hits == oracle;
      also with the MAG hashing sharded and the positions exchanged by bit range (all_to_all).
tests/test_multi_gpu.py launches this under `pytest -m gpu` on min(device_count, 4) GPUs.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist

from metasbt_b200 import ops, shard
from oracle import oracle as O
from tests import util


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ctx = ops.Context(local)
    k, m, n = 21, 1 << 20, 24
    w = ops.filter_words(m)
    fastas = [util.codes_to_fasta(util.mutate(util.synthetic_codes(100 + g // 4, 60_000), 0.01, g), "g%d" % g) for g in range(n)]
    want = [O.makebf(f, k, m) for f in fastas]

    # K1: this rank's block
    b, e = shard.genome_blocks([60_000] * n, world, k)[rank]
    batch = ops.GenomeBatch([ops.pack_fasta(f, k) for f in fastas[b:e]], dev)
    mine = ctx.build_filters(batch, k, m)
    torch.cuda.synchronize()
    for i in range(b, e):
        assert np.array_equal(mine[i - b, :w].cpu().numpy().view(np.uint64), want[i]), "K1 rank %d genome %d" % (rank, i)

    # K2: partial union per rank, merged by bit range over NCCL
    partial = ctx.bf_or([mine[i] for i in range(e - b)], m) if e > b else torch.zeros(ops.filter_stride_words(m), dtype=torch.int64, device=dev)
    merged = shard.or_merge(partial[:w].contiguous(), m, ctx=ctx)
    assert np.array_equal(merged.cpu().numpy().view(np.uint64), O.bf_or(want)), "K2 merge rank %d" % rank

    # replicate all leaves (small here) so that every rank can take its bit range of every leaf
    counts = [shard.genome_blocks([60_000] * n, world, k)[r] for r in range(world)]
    gathered = []
    for r, (rb, re_) in enumerate(counts):
        buf = mine[:, :w].contiguous() if r == rank else torch.empty((re_ - rb, w), dtype=torch.int64, device=dev)
        dist.broadcast(buf, src=r)
        gathered.append(buf)
    leaves_mat = torch.cat(gathered)
    leaves = [leaves_mat[i] for i in range(n)]
    rb, re_ = shard.bit_ranges(m, world)[rank]

    # K3: focus 0 against all, on this rank's bit range only (views of the rows), summed over ranks
    wb, we = rb // 64, (re_ + 63) // 64
    part_i, part_u = ctx.bf_dist(leaves_mat[0, wb:we].contiguous(), [leaves_mat[i, wb:we].contiguous() for i in range(n)], (we - wb) * 64)
    shard.all_reduce_counts(part_i)
    shard.all_reduce_counts(part_u)
    assert part_i.tolist() == [O.bf_and_popcount(want[0], x) for x in want], "K3 intersect rank %d" % rank
    assert part_u.tolist() == [O.bf_or_popcount(want[0], x) for x in want], "K3 union rank %d" % rank

    # K4: queries hashed on every rank, probed against this rank's rows only, partial hits summed
    queries = [util.codes_to_fasta(util.mutate(util.synthetic_codes(100 + s, 40_000), 0.04, 900 + s), "q%d" % s) for s in range(5)]
    qf = ctx.build_filters(ops.GenomeBatch([ops.pack_fasta(f, k) for f in queries], dev), k, m)
    positions = [ctx.extract_positions(qf[i], m) for i in range(len(queries))]
    sliced = ctx.bitslice(leaves, rb, re_)
    hits = ctx.query_batch_sliced(sliced, n, rb, re_, positions)
    shard.all_reduce_counts(hits)
    ref = np.stack([O.query_hits(O.positions_distinct(q, k, m), want) for q in queries])
    assert np.array_equal(hits.cpu().numpy().astype(np.uint64), ref), "K4 rank %d" % rank
    # the same through shard.query_sharded (each rank extracts only the positions of its own bit range)
    hits2 = shard.query_sharded(ctx, sliced, n, m, [qf[i] for i in range(len(queries))])
    assert np.array_equal(hits2.cpu().numpy().astype(np.uint64), ref), "K4 query_sharded rank %d" % rank
    # the same with the hashing sharded too: every rank hashes its block of the queries, positions are exchanged by bit range
    qb, qe = shard.query_blocks(len(queries), world)[rank]
    timings = dict()
    hits3 = shard.query_exchange_sharded(ctx, sliced, n, m, [qf[i] for i in range(qb, qe)], timings=timings)
    assert np.array_equal(hits3.cpu().numpy().astype(np.uint64), ref), "K4 query_exchange_sharded rank %d" % rank
    assert "ms_probe" in timings
    hits4 = shard.query_exchange_sharded(ctx, sliced, n, m, [qf[i] for i in range(qb, qe)], cap=(qe - qb) * 40_000, batch_size=len(queries))
    assert np.array_equal(hits4.cpu().numpy().astype(np.uint64), ref), "K4 query_exchange_sharded (cap, batch_size) rank %d" % rank
    # K3 all pairs over the filter set as it is spread over the ranks (half ring of rectangular launches over NCCL send/recv)
    local = shard.all_pairs_ring(ctx, mine, m, chunk=3)
    full = shard.all_pairs_assemble(local)
    want_pairs = np.array([[O.bf_and_popcount(a, b) for b in want] for a in want], dtype=np.int64)
    assert np.array_equal(full.cpu().numpy(), want_pairs), "K3 all_pairs_ring rank %d" % rank
    # the same matrix the K4 way (bit ranges of every filter exchanged, transposed and probed with their own set bits)
    probed = shard.all_pairs_probe(ctx, mine, m, chunk=2, batch=3)
    assert np.array_equal(probed.cpu().numpy(), want_pairs), "K3 all_pairs_probe rank %d" % rank
    dist.barrier()
    if rank == 0:
        print("multi-GPU parity ok: world=%d, K1 blocks %s, K2 OR-merge, K3 and K4 bit-range shards all match the oracle" % (world, counts))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
