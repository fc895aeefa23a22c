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
        open(os.path.join(tmp, "ok%d" % rank), "w").close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_partition_and_merge_under_gloo(tmp_path, oracle, world):
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / ("ok%d" % r)).exists() for r in range(world))
