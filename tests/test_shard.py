"""world_size-2 (and 3) gloo runs of the multi-GPU partition / merge logic on CPU:
genome blocks cover the batch exactly once; bit-range partial counts add up to the oracle's whole-filter
counts; the bit-range OR merge equals the oracle's OR."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from metasbt_b200 import shard


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_genome_blocks_cover_exactly_once():
    for world in (1, 2, 3, 8):
        for lens in ([5_000_000] * 1000, [10, 20, 30], [7], list(range(1, 50)), []):
            blocks = shard.genome_blocks(lens, world, k=31)
            assert len(blocks) == world
            flat = [i for b, e in blocks for i in range(b, e)]
            assert flat == list(range(len(lens)))
    b = shard.genome_blocks([5_000_000] * 1000, 8, k=31)
    assert max(e - s for s, e in b) - min(e - s for s, e in b) <= 1


def test_bit_ranges_are_aligned_partitions():
    for bits in (64, 1000, 1 << 22, (1 << 22) + 64, 1 << 30):
        for world in (1, 2, 3, 8):
            r = shard.bit_ranges(bits, world)
            assert r[0][0] == 0 and r[-1][1] == bits
            for (b0, e0), (b1, e1) in zip(r, r[1:]):
                assert e0 == b1 and b1 % 128 == 0


def _worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O
    from tests import util
    try:
        k, bits = 21, 1 << 16
        fastas = [util.codes_to_fasta(util.mutate(util.synthetic_codes(7, 9000), 0.02 * g, g), "g%d" % g) for g in range(7)]
        # K1: genome-sharded construction, no collective -- every rank builds its block
        blocks = shard.genome_blocks([9000] * len(fastas), world, k)
        b, e = blocks[rank]
        mine = {g: O.makebf(fastas[g], k, bits) for g in range(b, e)}
        full = [O.makebf(f, k, bits) for f in fastas]  # checker
        for g, w in mine.items():
            assert np.array_equal(w, full[g])
        # K2: OR-merge of per-rank partial unions by bit range
        partial = np.zeros(bits // 64, dtype=np.uint64)
        for w in mine.values():
            partial |= w
        merged = shard.or_merge(torch.from_numpy(partial.view(np.int64)), bits)
        assert np.array_equal(merged.numpy().view(np.uint64), O.bf_or(full))
        # K3: bit-range-sharded AND popcounts -> all_reduce(SUM)
        rb, re_ = shard.bit_ranges(bits, world)[rank]
        part = torch.tensor([O.bf_and_popcount(full[0][rb // 64: re_ // 64], f[rb // 64: re_ // 64]) for f in full], dtype=torch.int64)
        shard.all_reduce_counts(part)
        assert part.tolist() == [O.bf_and_popcount(full[0], f) for f in full]
        # K4: positions restricted to this rank's rows -> partial hits -> all_reduce(SUM)
        pos = O.positions_distinct(fastas[3], k, bits)
        sel = pos[(pos >= rb) & (pos < re_)]
        hits = torch.from_numpy(O.query_hits(sel, full).astype(np.int64))
        shard.all_reduce_counts(hits)
        assert hits.tolist() == [int(x) for x in O.query_hits(pos, full)]
        # K4 with the hashing sharded too (shard.query_exchange_sharded's host plan): every rank "hashes" its block of the
        # query batch, positions travel to the rank that owns their bit range, partial hits are summed
        queries = [fastas[g] for g in (3, 0, 5, 6, 1)]
        qb, qe = shard.query_blocks(len(queries), world)[rank]
        ranges = shard.bit_ranges(bits, world)
        local = [O.positions_distinct(q, k, bits) for q in queries[qb:qe]]
        counts = [[int(((p >= b0) & (p < e0)).sum()) for (b0, e0) in ranges] for p in local]
        gathered = [None] * world
        dist.all_gather_object(gathered, counts)
        send_sizes, recv_sizes, q_ranges = shard.exchange_plan(gathered, rank)
        send = [torch.from_numpy(np.concatenate([p[(p >= b0) & (p < e0)] - np.uint64(b0) for p in local] + [np.zeros(0, np.uint64)]).astype(np.int64))
                for (b0, e0) in ranges]
        assert [len(t) for t in send] == send_sizes
        recv = [torch.empty(n, dtype=torch.int64) for n in recv_sizes]
        shard._all_to_all_gloo(recv, send, None)
        flat = torch.cat(recv).numpy().astype(np.uint64)
        assert len(q_ranges) == len(queries) and q_ranges[-1][1] == len(flat)
        rows = [f[rb // 64: re_ // 64] for f in full]
        part = np.stack([O.query_hits(flat[b0:e0], rows) for b0, e0 in q_ranges]).astype(np.int64)
        hits = torch.from_numpy(part)
        shard.all_reduce_counts(hits)
        assert np.array_equal(hits.numpy().astype(np.uint64), np.stack([O.query_hits(O.positions_distinct(q, k, bits), full) for q in queries]))
        # K3 all pairs over a filter set spread across the ranks (half ring; every pair computed once)
        class OracleOps:  # stands in for ops.Context in this CPU test: the two launches all_pairs_ring makes
            def bf_allpairs(self, rows, num_bits):
                return torch.tensor([[O.bf_and_popcount(a.numpy().view(np.uint64), b.numpy().view(np.uint64)) for b in rows] for a in rows],
                                    dtype=torch.int64).reshape(len(rows), len(rows))

            def bf_rect(self, a_rows, b_rows, num_bits):
                return torch.tensor([[O.bf_and_popcount(a.numpy().view(np.uint64), b.numpy().view(np.uint64)) for b in b_rows] for a in a_rows],
                                    dtype=torch.int64).reshape(len(a_rows), len(b_rows))
            def bf_allpairs_probe(self, rows, num_bits, batch=256, slice_bytes=None):   # the K4 route's single launch site
                w = (num_bits + 63) // 64
                return self.bf_allpairs([r[:w] for r in rows], num_bits)
        for total in (7, 2 * world + 1, world):  # ragged blocks, one filter per rank
            lo, hi = total * rank // world, total * (rank + 1) // world
            pool = [O.makebf(util.codes_to_fasta(util.mutate(util.synthetic_codes(9, 5000), 0.03 * g, 70 + g), "p%d" % g), k, bits) for g in range(total)]
            mat = torch.from_numpy(np.stack(pool[lo:hi]).view(np.int64)) if hi > lo else torch.zeros((0, bits // 64), dtype=torch.int64)
            local = shard.all_pairs_ring(OracleOps(), mat, bits, chunk=2)
            full = shard.all_pairs_assemble(local)
            want = np.array([[O.bf_and_popcount(a, b) for b in pool] for a in pool], dtype=np.int64)
            assert np.array_equal(full.numpy(), want), (total, rank)
            # the same matrix the K4 way: bit ranges of every filter exchanged, partial matrices summed
            for odd_bits in (bits, bits - 37):
                pool2 = [p.copy() for p in pool]
                if odd_bits != bits:
                    for p in pool2:  # a filter of odd_bits bits: nothing set beyond it
                        p[odd_bits // 64] &= np.uint64((1 << (odd_bits % 64)) - 1)
                        p[odd_bits // 64 + 1:] = 0
                mat2 = torch.from_numpy(np.stack(pool2[lo:hi]).view(np.int64)) if hi > lo else torch.zeros((0, bits // 64), dtype=torch.int64)
                probed = shard.all_pairs_probe(OracleOps(), mat2, odd_bits, chunk=2)
                want2 = np.array([[O.bf_and_popcount(a, b) for b in pool2] for a in pool2], dtype=np.int64)
                assert np.array_equal(probed.numpy(), want2), (total, rank, odd_bits)
            computed = torch.tensor([int((local >= 0).sum())], dtype=torch.int64)
            dist.all_reduce(computed)
            assert int(computed) == sum((total * (r + 1) // world - total * r // world) ** 2 for r in range(world)) + \
                (total * total - sum((total * (r + 1) // world - total * r // world) ** 2 for r in range(world))) // 2  # each cross pair once
        open(os.path.join(tmp, "ok%d" % rank), "w").close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_partition_and_merge_under_gloo(tmp_path, oracle, world):
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / ("ok%d" % r)).exists() for r in range(world))


def test_query_blocks_and_exchange_plan():
    for n in (0, 1, 5, 256):
        for world in (1, 2, 3, 8):
            blocks = shard.query_blocks(n, world)
            assert [i for b, e in blocks for i in range(b, e)] == list(range(n))
    # two ranks, rank 0 hashed 2 MAGs, rank 1 hashed 1; counts[s][q][d]
    counts = [[[3, 4], [5, 0]], [[1, 2]]]
    send, recv, ranges = shard.exchange_plan(counts, 0)
    assert send == [8, 4] and recv == [8, 1] and ranges == [(0, 3), (3, 8), (8, 9)]
    send, recv, ranges = shard.exchange_plan(counts, 1)
    assert send == [1, 2] and recv == [4, 2] and ranges == [(0, 4), (4, 4), (4, 6)]
