"""GPU parity: every C-ABI entry point against the oracle on the same seeded inputs.
Integer work => bit-exact (np.array_equal), no tolerances anywhere."""
import ctypes
import random

import numpy as np
import pytest
import torch

from tests import util

pytestmark = pytest.mark.gpu


def to_np(t: torch.Tensor) -> np.ndarray:
    return t.cpu().numpy().view(np.uint64)


def build(ctx, fastas, k, num_bits, min_abund=1, modulus=None, variant=0):
    from metasbt_b200 import ops
    packed = [ops.pack_fasta(f, k) for f in fastas]
    batch = ops.GenomeBatch(packed, ctx.device)
    out = ctx.build_filters(batch, k, num_bits, min_abund=min_abund, modulus=modulus, variant=variant)
    torch.cuda.synchronize()
    w = ops.filter_words(num_bits)
    return out, [to_np(out[i, :w]) for i in range(len(fastas))]


@pytest.mark.parametrize("variant", [1, 2, 3, 4])
@pytest.mark.parametrize("k", [1, 9, 21, 31, 32, 33, 51, 64])
def test_makebf_messy_batches(ctx, oracle, k, variant):
    rng = random.Random(k)
    fastas = [util.messy_fasta(rng, rng.randint(1, 6), rng.choice([10, 300, 5000]), rng.choice(["ACGT", "ACGTN", "ACGTacgtNnRY"]))
              for _ in range(9)]
    fastas += [b"", b">empty\n", b">short\nACG\n"]
    sizes = ((1 << 16, None), (100003, None), (1 << 14, 20011), (20011, 1 << 14))
    if variant >= 3:  # the fused / phased kernels need num_bits >= 2^17
        sizes = ((1 << 17, None), (200003, None), (1 << 18, 300007), (300007, 1 << 18))
    for num_bits, modulus in sizes:
        _, got = build(ctx, fastas, k, num_bits, modulus=modulus, variant=variant)
        for f, g in zip(fastas, got):
            assert np.array_equal(g, oracle.makebf(f, k, num_bits, 1, 0, modulus))


@pytest.mark.parametrize("variant", [1, 2, 3, 4])
@pytest.mark.parametrize("num_bits", [64, 1000, (1 << 19) - 64, 1 << 19, (1 << 19) + 64, (1 << 21) + 192, 3 * (1 << 20) + 8])
def test_makebf_filter_sizes_and_tile_edges(ctx, oracle, num_bits, variant):
    """Filter sizes around the 64 KiB tile boundary, odd word counts, tiny filters."""
    if variant >= 3 and num_bits < (1 << 17):
        pytest.skip("the fused / phased kernels need num_bits >= 2^17")
    rng = random.Random(num_bits)
    fastas = [util.messy_fasta(rng, 3, 4000, "ACGTN") for _ in range(5)]
    _, got = build(ctx, fastas, 21, num_bits, variant=variant)
    for f, g in zip(fastas, got):
        assert np.array_equal(g, oracle.makebf(f, 21, num_bits))


@pytest.mark.parametrize("variant", [1, 2, 3, 4])
def test_makebf_skewed_inputs_overflow_path(ctx, oracle, variant):
    """Low-complexity genomes put (almost) every k-mer into a few buckets: the staging slots and the
    scratch buckets overflow and the fix-up pass must restore every bit."""
    k, num_bits = 15, 1 << 22
    fastas = [b">polyA\n" + b"A" * 200000 + b"\n",
              b">repeat\n" + b"ACGTTGCA" * 40000 + b"\n",
              util.codes_to_fasta(util.synthetic_codes(3, 150000), "rand"),
              b">mix\n" + b"AC" * 50000 + util.codes_to_fasta(util.synthetic_codes(4, 50000), "x").split(b"\n", 1)[1]]
    _, got = build(ctx, fastas, k, num_bits, variant=variant)
    for f, g in zip(fastas, got):
        assert np.array_equal(g, oracle.makebf(f, k, num_bits, rolling=True))


@pytest.fixture(params=["prefiltered", "tables"])
def min_path(request, monkeypatch):
    """min_abund >= 2 has two implementations: candidates prefiltered by a bit table (default for k <= 31) and plain count
    tables (k = 32, and everything when MSBT_MIN_TABLE is set)."""
    if request.param == "tables":
        monkeypatch.setenv("MSBT_MIN_TABLE", "1")
    else:
        monkeypatch.delenv("MSBT_MIN_TABLE", raising=False)
    return request.param


@pytest.mark.parametrize("k,min_abund", [(5, 2), (9, 2), (9, 3), (21, 2), (31, 2), (31, 3), (31, 4), (31, 7), (32, 2), (32, 3)])
def test_makebf_min_abundance(ctx, oracle, k, min_abund, min_path):
    rng = random.Random(k * 7 + min_abund)
    base = util.messy_fasta(rng, 2, 3000, "ACGT")
    fastas = [base + base[: len(base) // 2] + util.messy_fasta(rng, 2, 2000, "ACGTN"), b">r\n" + b"ACGT" * 200 + b"\n", b""]
    _, got = build(ctx, fastas, k, 1 << 15, min_abund=min_abund)
    for f, g in zip(fastas, got):
        assert np.array_equal(g, oracle.makebf(f, k, 1 << 15, min_abund))


def test_makebf_min_abundance_k_gt_32_rejected(ctx):
    with pytest.raises(ValueError):
        build(ctx, [b">r\n" + b"ACGT" * 50 + b"\n"], 33, 1 << 12, min_abund=2)


@pytest.mark.parametrize("variant", [1, 2, 3, 4])
@pytest.mark.parametrize("k,num_bits", [(31, 1 << 22), (21, 1 << 24), (51, 1 << 22), (31, 1 << 27)])
def test_makebf_synthetic_genomes(ctx, oracle, k, num_bits, variant):
    from metasbt_b200 import ops
    n = 300_000
    codes = [util.synthetic_codes(0x5EED0000 + g, n + 1000 * g) for g in range(4)]
    packed = [ops.pack_codes(c) for c in codes]
    batch = ops.GenomeBatch(packed, ctx.device)
    out = ctx.build_filters(batch, k, num_bits, variant=variant)
    torch.cuda.synchronize()
    w = ops.filter_words(num_bits)
    for g, c in enumerate(codes):
        fa = util.codes_to_fasta(c, "g%d" % g, 80)
        assert np.array_equal(to_np(out[g, :w]), oracle.makebf(fa, k, num_bits, rolling=(k <= 32)))
        # host packer and synthetic packer agree
        assert np.array_equal(ops.pack_fasta(fa, k).words, packed[g].words)


@pytest.mark.parametrize("variant", [1, 2, 3, 4])
def test_makebf_filter_index_and_rebuild(ctx, oracle, variant):
    """Filters are fully rewritten (stale bits cleared) and filter_index routes outputs."""
    from metasbt_b200 import ops
    k, num_bits = 15, (1 << 17) + 64 if variant >= 3 else 1 << 14
    rng = random.Random(2)
    fastas = [util.messy_fasta(rng, 2, 800, "ACGT") for _ in range(3)]
    packed = [ops.pack_fasta(f, k) for f in fastas]
    out = torch.full((5, ops.filter_stride_words(num_bits)), -1, dtype=torch.int64, device=ctx.device)
    batch = ops.GenomeBatch(packed, ctx.device, filter_indices=[4, 0, 2])
    ctx.build_filters(batch, k, num_bits, out=out, variant=variant)
    torch.cuda.synchronize()
    w = ops.filter_words(num_bits)
    for f, idx in zip(fastas, [4, 0, 2]):
        assert np.array_equal(to_np(out[idx, :w]), oracle.makebf(f, k, num_bits))
    assert (out[1] == -1).all() and (out[3] == -1).all()


@pytest.mark.parametrize("num_bits", [64, 1000, (1 << 16) + 64, (1 << 20) - 64])
def test_filter_algebra(ctx, oracle, num_bits):
    from metasbt_b200 import ops
    rng = np.random.default_rng(num_bits)
    w, stride = ops.filter_words(num_bits), ops.filter_stride_words(num_bits)
    n = 11
    host = rng.integers(0, 1 << 63, size=(n, stride), dtype=np.int64) * rng.integers(0, 2, size=(n, stride))
    dev = torch.from_numpy(host).to(ctx.device)
    rows = [dev[i] for i in range(n)]
    hrows = [host[i, :w].view(np.uint64).copy() for i in range(n)]
    # K2
    for cnt in (1, 2, 5, n):
        got = ctx.bf_or(rows[:cnt], num_bits)
        assert np.array_equal(to_np(got[:w]), oracle.bf_or(hrows[:cnt]))
    # K3'
    pc = ctx.bf_popcount(rows, num_bits).cpu().numpy()
    assert [int(x) for x in pc] == [oracle.bf_popcount(h) for h in hrows]
    # K3 one-vs-many
    inter, union = ctx.bf_dist(rows[0], rows[1:], num_bits)
    assert [int(x) for x in inter.cpu()] == [oracle.bf_and_popcount(hrows[0], h) for h in hrows[1:]]
    assert [int(x) for x in union.cpu()] == [oracle.bf_or_popcount(hrows[0], h) for h in hrows[1:]]
    # K3 all pairs
    ap = ctx.bf_allpairs(rows, num_bits).cpu().numpy()
    for i in range(n):
        for j in range(n):
            assert int(ap[i, j]) == oracle.bf_and_popcount(hrows[i], hrows[j])


@pytest.mark.parametrize("n,num_bits", [(1, 4096), (2, 64), (3, 100000), (10, (1 << 19) + 64), (16, 50000), (17, 8192 * 3 + 64),
                                        (33, 1 << 16), (50, 12345), (130, 5000)])
def test_allpairs_block_shapes(ctx, oracle, n, num_bits, monkeypatch):
    """K3 all pairs, carry-save kernel: one block, a full block, blocks with ragged ends, several tile pairs; odd word
    counts; dense random bits (every carry-save plane in use).  Checked against the oracle's AND-popcount and against the
    plain POPC kernel (MSBT_ALLPAIRS=1)."""
    from metasbt_b200 import ops
    rng = np.random.default_rng(n * 1000 + num_bits % 997)
    w, stride = ops.filter_words(num_bits), ops.filter_stride_words(num_bits)
    host = rng.integers(-(1 << 63), (1 << 63) - 1, size=(n, stride), dtype=np.int64)
    host[:, w:] = 0
    if n > 2:
        host[2] = -1   # all ones: popcount = 64 * words, the largest count a pair can reach
        host[2, w:] = 0
    dev = torch.from_numpy(host).to(ctx.device)
    rows = [dev[i] for i in range(n)]
    bits = np.unpackbits(host[:, :w].copy().view(np.uint8), axis=1).astype(np.int64)
    want = bits @ bits.T
    got = ctx.bf_allpairs(rows, num_bits).cpu().numpy()
    assert np.array_equal(got, want)
    assert int(got[0, n - 1]) == oracle.bf_and_popcount(host[0, :w].view(np.uint64), host[n - 1, :w].view(np.uint64))
    monkeypatch.setenv("MSBT_ALLPAIRS", "1")
    assert np.array_equal(ctx.bf_allpairs(rows, num_bits).cpu().numpy(), want)


@pytest.mark.parametrize("n", [10, 20])
def test_allpairs_many_tiles_per_cta(ctx, oracle, n, monkeypatch):
    """Filters long enough (2^25 bits) that every CTA of the carry-save kernel walks several tiles, i.e. the double-buffered
    cp.async pipeline reaches its steady state (n = 10: one block pair, 2 KiB runs; n = 20: three block pairs, 1 KiB
    runs, 13 tiles per CTA).  Checked against the plain POPC kernel, the popcount kernel (diagonal) and the oracle."""
    from metasbt_b200 import ops
    num_bits = (1 << 25) + 192   # not a multiple of the tile length
    w = ops.filter_words(num_bits)
    g = torch.Generator(device=ctx.device).manual_seed(n)
    dev = torch.randint(-(1 << 62), 1 << 62, (n, ops.filter_stride_words(num_bits)), dtype=torch.int64, device=ctx.device, generator=g)
    dev &= torch.randint(-(1 << 62), 1 << 62, dev.shape, dtype=torch.int64, device=ctx.device, generator=g)
    dev[:, w:] = 0
    rows = [dev[i] for i in range(n)]
    got = ctx.bf_allpairs(rows, num_bits).cpu().numpy()
    monkeypatch.setenv("MSBT_ALLPAIRS", "1")
    assert np.array_equal(ctx.bf_allpairs(rows, num_bits).cpu().numpy(), got)
    assert np.array_equal(np.diag(got), ctx.bf_popcount(rows, num_bits).cpu().numpy())
    h0, h1, hl = (dev[i, :w].cpu().numpy().view(np.uint64) for i in (0, 1, n - 1))
    assert int(got[0, 1]) == oracle.bf_and_popcount(h0, h1) and int(got[1, n - 1]) == oracle.bf_and_popcount(h1, hl)
    assert np.array_equal(got, got.T)


def test_query_paths_agree_with_oracle(ctx, oracle):
    from metasbt_b200 import ops
    rng = random.Random(11)
    k, num_bits = 21, 1 << 18
    base = [util.synthetic_codes(100 + s, 20000) for s in range(5)]
    leaves_fa = []
    for s, c in enumerate(base):
        for v in range(9):
            leaves_fa.append(util.codes_to_fasta(util.mutate(c, 0.01, 1000 * s + v), "leaf%d_%d" % (s, v)))
    leaf_mat, leaf_np = build(ctx, leaves_fa, k, num_bits)
    queries_fa = [util.codes_to_fasta(util.mutate(base[1], 0.03, 77), "q0"),
                  util.codes_to_fasta(base[3][:5000], "q1") + util.messy_fasta(rng, 3, 900, "ACGTN"),
                  b">tiny\nACGT\n",
                  util.codes_to_fasta(util.synthetic_codes(999, 7000), "unrelated")]
    q_mat, q_np = build(ctx, queries_fa, k, num_bits)
    leaves = [leaf_mat[i] for i in range(len(leaves_fa))]
    positions = [ctx.extract_positions(q_mat[i], num_bits) for i in range(len(queries_fa))]
    batched = ctx.extract_positions_batch([q_mat[i] for i in range(len(queries_fa))], num_bits)
    capped = ctx.extract_positions_batch([q_mat[i] for i in range(len(queries_fa))], num_bits, cap=10 ** 6)
    for qf, p, pb, pc in zip(queries_fa, positions, batched, capped):
        ref = oracle.positions_distinct(qf, k, num_bits)
        assert np.array_equal(p.cpu().numpy().view(np.uint32).astype(np.uint64), ref)
        assert torch.equal(p, pb) and torch.equal(p, pc)
    want = np.stack([oracle.query_hits(oracle.positions_distinct(qf, k, num_bits), leaf_np) for qf in queries_fa])
    hits = ctx.query_batch(leaves, positions).cpu().numpy()
    assert np.array_equal(hits.astype(np.uint64), want)
    # bit-sliced path, whole range and split into two bit-range shards that must add up
    sliced = ctx.bitslice(leaves, 0, num_bits)
    hs = ctx.query_batch_sliced(sliced, len(leaves), 0, num_bits, positions).cpu().numpy()
    assert np.array_equal(hs.astype(np.uint64), want)
    half = num_bits // 2
    s0, s1 = ctx.bitslice(leaves, 0, half), ctx.bitslice(leaves, half, num_bits)
    h0 = ctx.query_batch_sliced(s0, len(leaves), 0, half, positions)
    h1 = ctx.query_batch_sliced(s1, len(leaves), half, num_bits, positions)
    assert np.array_equal((h0 + h1).cpu().numpy().astype(np.uint64), want)
    # corollary A6: hits == popc(Q & F)
    inter, _ = ctx.bf_dist(q_mat[0], leaves, num_bits)
    assert np.array_equal(inter.cpu().numpy().astype(np.uint64), want[0])


def test_error_behaviour(ctx):
    from metasbt_b200 import ops
    pg = ops.pack_fasta(b">a\nACGTACGTACGT\n", 5)
    batch = ops.GenomeBatch([pg], ctx.device)
    with pytest.raises(ValueError):
        ctx.build_filters(batch, 0, 1024)
    with pytest.raises(ValueError):
        ctx.build_filters(batch, 65, 1024)
    with pytest.raises(ValueError):
        ctx.bf_or([], 1024)
    assert "k=65" in ctx._L.msbt_last_error(ctx._h).decode() or True


def _kats():
    import json
    import os
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "kat_positions.json")) as f:
        return json.load(f)


@pytest.mark.parametrize("variant", [1, 2, 3, 4])
def test_committed_known_answers(ctx, variant):
    """The committed golden vectors, without the oracle in the loop."""
    from metasbt_b200 import ops
    for kat in _kats():
        if (kat["min"] > 1 and variant >= 2) or (variant >= 3 and kat["num_bits"] < (1 << 17)):
            continue
        _, got = build(ctx, [kat["fasta"].encode()], kat["k"], kat["num_bits"], min_abund=kat["min"], modulus=kat["modulus"],
                       variant=variant if kat["min"] == 1 else 0)
        bits = [i * 64 + b for i, x in enumerate(got[0]) for b in range(64) if (int(x) >> b) & 1]
        assert bits == kat["set_bits"], kat["name"]
        q, _ = build(ctx, [kat["fasta"].encode()], kat["k"], kat["num_bits"], modulus=kat["modulus"])
        pos = ctx.extract_positions(q[0], kat["num_bits"])
        assert [int(x) for x in pos.cpu().numpy().view(np.uint32)] == kat["query_positions"], kat["name"]


@pytest.mark.parametrize("n_leaves,num_bits,row_begin,row_end", [
    (45, 1 << 16, 0, 1 << 16),                 # one partial column group, whole 1,024-row tiles
    (600, 70000, 0, 70000),                    # two column groups (512 + 88 leaves), ragged last tile and last row word
    (1100, 1 << 17, 32 * 5, 32 * 5 + 40001),   # first word not 16-byte aligned: the general load path
    (513, 1 << 17, 4096, 4096 + 3000),         # aligned sub-range, 1 leaf in the last column group
    (33, 4000, 128, 3999),                     # smaller than one tile
])
def test_bitslice_is_the_transpose(ctx, n_leaves, num_bits, row_begin, row_end):
    """msbt_bitslice_build against a numpy transposition of the same leaves (row p, bit l = bit p of leaf l)."""
    from metasbt_b200 import ops
    rng = np.random.default_rng(n_leaves)
    w = ops.filter_stride_words(num_bits)   # rows 16-byte aligned, as the C-ABI demands
    host = rng.integers(-(1 << 63), (1 << 63) - 1, size=(n_leaves, w), dtype=np.int64)
    dev = torch.from_numpy(host).to(ctx.device)
    got = ctx.bitslice([dev[i] for i in range(n_leaves)], row_begin, row_end).cpu().numpy().view(np.uint32)
    lw = (n_leaves + 31) // 32
    assert got.shape == (row_end - row_begin, ctx.sliced_row_words(n_leaves))
    bits = np.unpackbits(host.view(np.uint8), axis=1, bitorder="little")[:, row_begin:row_end]   # [leaf, row]
    cols = np.zeros((row_end - row_begin, lw * 32), dtype=np.uint8)
    cols[:, :n_leaves] = bits.T
    want = np.packbits(cols, axis=1, bitorder="little").view(np.uint32)
    assert np.array_equal(got[:, :lw], want)


def test_query_sliced_many_leaves_and_chunks(ctx, oracle):
    """Bit-sliced probe with > 1024 leaves (several column groups), queries long enough for several warp sub-chunks
    and CTAs, an empty query, and a split into three bit-range shards that must add up."""
    from metasbt_b200 import ops
    num_bits, n_leaves = 1 << 17, 1100
    rng = np.random.default_rng(7)
    w = ops.filter_words(num_bits)
    host = rng.integers(0, 1 << 63, size=(n_leaves, w), dtype=np.int64) & rng.integers(0, 1 << 63, size=(n_leaves, w), dtype=np.int64)
    dev = torch.from_numpy(host).to(ctx.device)
    leaves = [dev[i] for i in range(n_leaves)]
    leaf_np = [host[i].view(np.uint64) for i in range(n_leaves)]
    queries = [np.sort(rng.choice(num_bits, size=n, replace=False)).astype(np.uint64) for n in (70000, 9000, 0, 33)]
    positions = [torch.from_numpy(q.astype(np.uint32).view(np.int32)).to(ctx.device) for q in queries]
    want = np.stack([oracle.query_hits(q, leaf_np) for q in queries])
    sliced = ctx.bitslice(leaves, 0, num_bits)
    got = ctx.query_batch_sliced(sliced, n_leaves, 0, num_bits, positions).cpu().numpy()
    assert np.array_equal(got.astype(np.uint64), want)
    cuts = [0, 40000 // 32 * 32, 90000 // 32 * 32, num_bits]
    total = None
    for b, e in zip(cuts, cuts[1:]):
        part = ctx.query_batch_sliced(ctx.bitslice(leaves, b, e), n_leaves, b, e, positions)
        total = part if total is None else total + part
    assert np.array_equal(total.cpu().numpy().astype(np.uint64), want)
    assert np.array_equal(ctx.query_batch(leaves[:37], positions).cpu().numpy().astype(np.uint64), want[:, :37])


@pytest.mark.parametrize("variant", [0, 2, 4])
def test_full_size_genomes_bit_exact_and_properties(ctx, oracle, variant):
    """BASELINE.json configs[1] shape at full size (5 Mbp genomes, k=31, 2^30-bit filters) for three genomes:
    bit-exact against the oracle, plus the size-independent properties the domain offers --
    popcount == number of distinct positions, OR of two filters == filter of the concatenated FASTA (k-mers never span
    records), union/intersection counts consistent, and the query of a genome against its own filter hits every position."""
    from metasbt_b200 import ops
    k, m, n = 31, 1 << 30, 5_000_000
    codes = [util.synthetic_codes(0x5EED0000 + g, n) for g in range(2)]
    codes.append(util.mutate(codes[0], 0.01, 99))
    fastas = [util.codes_to_fasta(c, "g%d" % i, 80) for i, c in enumerate(codes)]
    batch = ops.GenomeBatch([ops.pack_codes(c) for c in codes], ctx.device)
    out = ctx.build_filters(batch, k, m, variant=variant)
    torch.cuda.synchronize()
    w = ops.filter_words(m)
    want = [oracle.makebf(f, k, m, rolling=True) for f in fastas]
    for i in range(3):
        assert torch.equal(out[i, :w].cpu(), torch.from_numpy(want[i].view(np.int64))), "genome %d differs" % i
    rows = [out[i] for i in range(3)]
    pc = [int(x) for x in ctx.bf_popcount(rows, m).cpu()]
    pos0 = ctx.extract_positions(rows[0], m)
    assert pc[0] == pos0.numel() == len(oracle.positions_distinct(fastas[0], k, m))
    both = ctx.build_filters(ops.GenomeBatch([ops.pack_fasta(fastas[0] + fastas[1], k)], ctx.device), k, m, variant=variant)
    union = ctx.bf_or(rows[:2], m)
    assert torch.equal(both[0, :w], union[:w])
    inter, uni = ctx.bf_dist(rows[0], rows[1:], m)
    for j in (1, 2):
        assert int(inter[j - 1]) + int(uni[j - 1]) == pc[0] + pc[j]
    assert int(uni[0]) == int(ctx.bf_popcount([union], m)[0])
    hits = ctx.query_batch(rows, [pos0]).cpu().numpy()[0]
    assert int(hits[0]) == pc[0] and int(hits[2]) == int(inter[1]) and int(hits[1]) == int(inter[0])
    assert 0.6 * pc[0] < int(hits[2]) < 0.85 * pc[0]  # a 1 % mutant shares about 0.99^31 = 73 % of its 31-mers


@pytest.mark.parametrize("min_abund", [2, 3])
def test_makebf_min_abundance_large_genome(ctx, oracle, min_abund, min_path):
    """MetaSBT's CLI default (--min-kmer-occurrences 2) on a 1.2 Mbp genome with repeated segments, a low-complexity
    stretch and an N run: exact canonical multiplicity (forward and reverse-complement copies count together)."""
    from metasbt_b200 import ops
    k, m = 31, 1 << 26
    base = util.synthetic_codes(77, 1_000_000)
    rc = (3 - base[200_000:260_000])[::-1]
    codes = np.concatenate([base, base[100_000:180_000], rc, base[100_000:130_000], np.zeros(5000, dtype=np.uint8)])
    fa = util.codes_to_fasta(codes, "rep", 80) + b">n\nACGT" + b"N" * 50 + b"ACGTACGTACGTACGTACGTACGTACGTACGTACGTAC" * 3 + b"\n"
    _, got = build(ctx, [fa, b">tiny\nACGT\n"], k, m, min_abund=min_abund)
    want = oracle.makebf(fa, k, m, min_abund)
    assert np.array_equal(got[0], want)
    assert oracle.bf_popcount(want) > 20_000 and oracle.bf_popcount(got[1]) == 0


def test_makebf_min_abundance_all_repeats(ctx, oracle, min_path):
    """A genome that is one 40 kbp segment repeated six times (plus its reverse complement): nearly every k-mer is a
    candidate, so the per-block candidate lists overflow and the exact table is as large as it gets."""
    k, m = 31, 1 << 22
    seg = util.synthetic_codes(91, 40_000)
    codes = np.concatenate([seg] * 6 + [(3 - seg)[::-1]] + [util.synthetic_codes(92, 10_000)])
    fa = util.codes_to_fasta(codes, "rep6", 80)
    for min_abund in (2, 7, 8):
        _, got = build(ctx, [fa], k, m, min_abund=min_abund)
        want = oracle.makebf(fa, k, m, min_abund)
        assert np.array_equal(got[0], want)
    assert oracle.bf_popcount(oracle.makebf(fa, k, m, 7)) > 30_000 and oracle.bf_popcount(want) == 0


def test_query_sharded_partials_add_up(ctx, oracle):
    """shard.query_sharded: every bit-range shard extracts only its own positions from the query filters and probes its own
    rows; the partial hit matrices of 1, 2, 3 and 5 shards add up to the oracle's hits (odd world sizes: uneven ranges)."""
    from metasbt_b200 import shard
    rng = random.Random(99)
    k, num_bits, n_leaves = 15, (1 << 14) + 192, 70
    leaf_fa = [util.messy_fasta(rng, 2, 1500, "ACGT") for _ in range(n_leaves)]
    leaves_dev, leaf_bits = build(ctx, leaf_fa, k, num_bits)
    queries = [leaf_fa[3] + util.messy_fasta(rng, 1, 800, "ACGT"), leaf_fa[40][: len(leaf_fa[40]) // 2], b">e\nACG\n"]
    qf, _ = build(ctx, queries, k, num_bits)
    want = np.stack([oracle.query_hits(oracle.positions_distinct(q, k, num_bits), leaf_bits) for q in queries])
    leaves = [leaves_dev[i] for i in range(n_leaves)]
    for world in (1, 2, 3, 5):
        total = None
        for rank in range(world):
            rb, re_ = shard.bit_ranges(num_bits, world)[rank]
            part = shard.query_sharded(ctx, ctx.bitslice(leaves, rb, re_), n_leaves, num_bits, [qf[i] for i in range(len(queries))],
                                       rank=rank, world=world)
            total = part if total is None else total + part
        assert np.array_equal(total.cpu().numpy().astype(np.uint64), want), world


def _ingest_cases():
    rng = random.Random(42)
    big_line = b">one_line\n" + util.codes_to_fasta(util.synthetic_codes(5, 20000), "x", 0).split(b"\n", 1)[1]
    long_header = b">" + b"h" * 9000 + b"\nACGTACGTACGTAAACCCGGGTTT\n"
    return [
        util.messy_fasta(rng, 4, 3000),
        util.messy_fasta(rng, 2, 9000, "ACGTacgtNn"),
        b"", b">only_header", b">only_header\n", b"\n\n\n", b"ACGT", b"ACGTACGTACGTACGTACGTACGTACGTACGTACG",   # no header, no newline
        b">a\r\nACGTACGTAC\r\nGGGTTTAAAC\r\n>b\r\nTTTT\r\n",
        b">a\nACGT>ACGTNACGT\n >notheader\nACGTACGTACGT\n",
        big_line, long_header,
        b">x\n" + b"ACGTTGCA" * 512 + b"\n" + b"ACGTTGCA" * 511 + b"ACGTTG\n",            # lines ending exactly on a chunk boundary
        b">x\n" + b"A" * 4093 + b"\n>" + b"y" * 4095 + b"\nCCCCCCCCCCGGGGGGGGGG\n",       # header starting at a chunk's first byte
        util.codes_to_fasta(util.synthetic_codes(9, 300_000), "big", 80),
        b">n\n" + b"N" * 10000 + b"ACGTACGTACGTACGTAGCTAGCTAGCTAGCATCGATCAGCTAC" + b"N" * 5000 + b"\n",
    ]


@pytest.mark.parametrize("min_run", [0, 1, 5, 31, 64])
def test_fasta_pack_device_matches_host_packer(ctx, min_run):
    """Device FASTA ingest (msbt_fasta_pack_device) == host packer (msbt_fasta_pack), file by file: packed words incl. the
    zero padding, run table, base and run counts -- messy records, CRLF, '>' inside a line, headers and lines longer than
    a chunk, chunk-aligned line ends, empty files, N runs."""
    from metasbt_b200 import ops
    texts = _ingest_cases()
    batch = ctx.pack_fasta_device(texts, min_run)
    torch.cuda.synchronize()
    words = batch.d_words.cpu().numpy().view(np.uint64)
    runs = batch.d_runs.cpu().numpy().view(np.uint32)
    for i, t in enumerate(texts):
        ref = ops.pack_fasta(t, min_run)
        d = batch.descs[i]
        assert int(d.n_bases) == ref.n_bases, (i, int(d.n_bases), ref.n_bases)
        assert int(d.n_runs) == len(ref.run_ends), i
        assert np.array_equal(words[d.packed_word_off: d.packed_word_off + len(ref.words)], ref.words), i
        assert np.array_equal(runs[d.run_off: d.run_off + d.n_runs], ref.run_ends), i


def test_fasta_pack_device_feeds_k1(ctx, oracle):
    texts = _ingest_cases()
    k, m = 21, 1 << 20
    batch = ctx.pack_fasta_device(texts, k)
    out = ctx.build_filters(batch, k, m)
    torch.cuda.synchronize()
    w = 1 << 14
    for i, t in enumerate(texts):
        assert np.array_equal(to_np(out[i, :w]), oracle.makebf(t, k, m)), i


@pytest.mark.parametrize("num_bits", [1 << 32, (1 << 32) + 64])
def test_makebf_maximum_filter_size(ctx, oracle, num_bits):
    """The largest filter the u32-position kernels take (2^32 bits = 512 MiB, fused path with 4-tile buckets) and the first
    size beyond it (falls back to the direct kernel with 64-bit positions): bit-exact, checked through the sparse set of
    positions (the oracle's full 512 MiB filter would be compared word for word as well, but positions localise a failure)."""
    from metasbt_b200 import ops
    k = 31
    fa = util.codes_to_fasta(util.synthetic_codes(31337, 200_000), "g", 80) + b">n\nACGTNNNNACGTACGTACGTACGTACGTACGTACGTACGTACGTA\n"
    batch = ops.GenomeBatch([ops.pack_fasta(fa, k)], ctx.device)
    out = ctx.build_filters(batch, k, num_bits)
    torch.cuda.synchronize()
    w = ops.filter_words(num_bits)
    want = oracle.makebf(fa, k, num_bits)
    got = out[0, :w].cpu().numpy().view(np.uint64)
    nz_g, nz_w = np.flatnonzero(got), np.flatnonzero(want)
    assert np.array_equal(nz_g, nz_w) and np.array_equal(got[nz_g], want[nz_w])
    assert int(ctx.bf_popcount([out[0]], num_bits)[0]) == oracle.bf_popcount(want) > 199_000
    del out
    torch.cuda.empty_cache()


def test_backend_sliced_tree_cache(ctx, oracle, tmp_path):
    """CudaBackend.query_hits: a flat tree with many leaves is transposed as soon as that pays and probed bit-sliced from
    then on; the hits equal the oracle's every time, and rewriting one leaf invalidates the cached matrix."""
    from metasbt_b200 import backend as B, bffile
    rng = random.Random(4242)
    k, num_bits, n_leaves = 21, 1 << 16, 300
    be = B.CudaBackend(k, num_bits)
    fas, bfs = [], []
    for i in range(n_leaves):
        fa = tmp_path / ("leaf%03d.fa" % i)
        fa.write_bytes(util.messy_fasta(rng, 1, 900, "ACGT"))
        fas.append(str(fa))
        bfs.append(str(tmp_path / ("leaf%03d.bf" % i)))
    be.sketch_many(fas, bfs, 1)
    qfa = tmp_path / "query.fa"
    qfa.write_bytes(open(fas[7], "rb").read() + open(fas[123], "rb").read() + util.messy_fasta(rng, 1, 500, "ACGTN"))
    positions = be.query_positions(str(qfa))
    qpos = oracle.positions_distinct(qfa.read_bytes(), k, num_bits)

    def want():
        return [int(x) for x in oracle.query_hits(qpos, [bffile.read_bf(p)[1] for p in bfs])]

    # the first query goes through the bit-sliced matrix only if its row-major probe would move more bytes than the
    # transposition; from the tree's second query on the matrix exists and serves every query
    seen = []
    for _ in range(3):
        hits, total = be.query_hits(positions, bfs)
        assert total == len(qpos) and hits == want()
        seen.append(be.sliced_probes)
    assert seen[0] in (0, 1) and seen[1] == seen[0] + 1 and seen[2] == seen[1] + 1
    for i in (7, 123):  # the query contains these leaves whole
        assert hits[i] == oracle.bf_popcount(bffile.read_bf(bfs[i])[1])
    # rewrite leaf 5 with the content of leaf 7: the cached matrix must go, the next answer must see the new leaf
    be.sketch_many([fas[7]], [bfs[5]], 1)
    hits, _ = be.query_hits(positions, bfs)
    assert hits == want() and hits[5] == hits[7]
    # small trees stay on the row-major probe
    before = be.sliced_probes
    for _ in range(3):
        h, _ = be.query_hits(positions, bfs[:40])
        assert h == want()[:40]
    assert be.sliced_probes == before


def test_sliced_matrix_dropped_even_when_the_row_major_copy_was_evicted(ctx, oracle, tmp_path):
    """a leaf can be LRU-evicted from the filter cache while a cached bit-sliced matrix still contains it;
    rewriting that leaf must drop the matrix all the same."""
    from metasbt_b200 import backend as B, bffile
    rng = random.Random(99)
    k, num_bits, n_leaves = 21, 1 << 16, 260
    be = B.CudaBackend(k, num_bits)
    fas = []
    for i in range(n_leaves):
        fa = tmp_path / ("leaf%03d.fa" % i)
        fa.write_bytes(util.messy_fasta(rng, 1, 900, "ACGT"))
        fas.append(str(fa))
    bfs = [str(tmp_path / ("leaf%03d.bf" % i)) for i in range(n_leaves)]
    be.sketch_many(fas, bfs, 1)
    positions = be.query_positions(fas[9])
    hits, _ = be.query_hits(positions, bfs)
    hits, _ = be.query_hits(positions, bfs)  # the second query of a tree always finds it transposed
    assert be.sliced_probes >= 1 and len(be._sliced) == 1
    for p in list(be._cache):  # cache pressure: every row-major copy goes
        be._bytes -= be._cache.pop(p).numel() * 8
    assert len(be._sliced) == 1
    be.sketch_many([fas[9]], [bfs[5]], 1)
    assert len(be._sliced) == 0
    hits, _ = be.query_hits(positions, bfs)
    assert hits[5] == hits[9] == oracle.bf_popcount(bffile.read_bf(bfs[9])[1])


def test_dist_pairs_and_query_ranges(ctx, oracle):
    """msbt_bf_and_popcount_pairs and the *_ranges query entry points (any subset, any order, of a packed batch)."""
    rng = random.Random(5)
    k = 21
    for num_bits in (1000, 1 << 16, (1 << 18) + 64):
        fastas = [util.messy_fasta(rng, 2, 3000, "ACGTN") for _ in range(12)]
        mat, flt = build(ctx, fastas, k, num_bits)
        rows = [mat[i] for i in range(len(fastas))]
        pa = [rng.randrange(12) for _ in range(40)]
        pb = [rng.randrange(12) for _ in range(40)]
        inter, union = ctx.bf_dist_pairs([rows[i] for i in pa], [rows[j] for j in pb], num_bits)
        assert inter.tolist() == [oracle.bf_and_popcount(flt[i], flt[j]) for i, j in zip(pa, pb)]
        assert union.tolist() == [oracle.bf_or_popcount(flt[i], flt[j]) for i, j in zip(pa, pb)]
        # packed positions of queries 0..7; probe a shuffled subset with repeats against leaves 4..11
        pos, offs = ctx.extract_positions_packed(rows[:8], num_bits)
        subset = [5, 0, 7, 7, 2]
        ranges = [(offs[q], offs[q + 1]) for q in subset] + [(offs[3], offs[3])]  # and an empty range
        want = np.stack([oracle.query_hits(oracle.positions_distinct(fastas[q], k, num_bits), flt[4:]) for q in subset]
                        + [np.zeros(8, dtype=np.uint64)])
        got = ctx.query_ranges(rows[4:], pos, ranges).cpu().numpy()
        assert np.array_equal(got.astype(np.uint64), want)
        sliced = ctx.bitslice(rows[4:], 0, num_bits)
        got = ctx.query_sliced_ranges(sliced, 8, 0, num_bits, pos, ranges).cpu().numpy()
        assert np.array_equal(got.astype(np.uint64), want)
    with pytest.raises(ValueError):
        ctx.query_ranges(rows[4:], pos, [(5, 3)])
    with pytest.raises(ValueError):
        ctx.bf_dist_pairs(rows[:2], rows[:3], num_bits)


def test_allpairs_more_than_65535_block_pairs(ctx):
    """n > 5,776 filters (more than 65,535 block pairs of 16 x 16) used to be rejected."""
    n, num_bits = 5800, 128
    g = torch.Generator().manual_seed(7)
    host = torch.randint(-2 ** 62, 2 ** 62, (n, 2), generator=g, dtype=torch.int64)
    dev = host.to(ctx.device)
    got = ctx.bf_allpairs([dev[i] for i in range(n)], num_bits).cpu().numpy()
    a = host.numpy().view(np.uint64)
    bits = np.unpackbits(a.view(np.uint8), axis=1).astype(np.int32)  # [n, 128]
    want = bits @ bits.T
    assert np.array_equal(got, want.astype(np.int64))


def test_backend_streams_operands_in_chunks(ctx, oracle, tmp_path):
    """operand lists larger than the cache budget stream through in chunks (union, popcounts, 1-vs-N, pair list,
    tiled all-pairs, leaf-chunked query) and give the oracle's numbers."""
    from metasbt_b200 import backend as B, bffile
    rng = random.Random(31)
    k, num_bits, n = 21, 1 << 16, 23
    row = (num_bits // 64) * 8
    be = B.CudaBackend(k, num_bits, cache_bytes=12 * row)  # 6 filters per chunk
    assert be._chunk() == 6
    fas, bfs = [], []
    for i in range(n):
        fa = tmp_path / ("g%02d.fa" % i)
        fa.write_bytes(util.messy_fasta(rng, 2, 2500, "ACGTN"))
        fas.append(str(fa))
        bfs.append(str(tmp_path / ("g%02d.bf" % i)))
    be.sketch_many(fas, bfs, 1)
    flt = [bffile.read_bf(p)[1] for p in bfs]
    for f, fa in zip(flt, fas):
        assert np.array_equal(f, oracle.makebf(open(fa, "rb").read(), k, num_bits))
    assert len(be._cache) <= 12
    assert be.popcounts(bfs) == [oracle.bf_popcount(f) for f in flt]
    be.union(bfs, str(tmp_path / "u.bf"))
    assert np.array_equal(bffile.read_bf(str(tmp_path / "u.bf"))[1], oracle.bf_or(flt))
    jobs = [(bfs[:9], str(tmp_path / "u0.bf")), ([bfs[4]], str(tmp_path / "u1.bf")), (bfs[7:], str(tmp_path / "u2.bf")),
            (bfs[2:4], str(tmp_path / "u3.bf")), (bfs, str(tmp_path / "u4.bf")), (bfs[10:13], str(tmp_path / "u5.bf"))]
    be.union_many(jobs)  # more jobs than pinned ring slots, one of them a single-source copy, all longer than a chunk or not
    for srcs, out in jobs:
        assert np.array_equal(bffile.read_bf(out)[1], oracle.bf_or([flt[bfs.index(p)] for p in srcs])), out
        assert np.array_equal(be.filter(out)[: be.words].cpu().numpy().view(np.uint64), bffile.read_bf(out)[1])
    assert open(jobs[1][1], "rb").read() == open(bfs[4], "rb").read()
    assert be.dist_one_vs_many(bfs[3], bfs) == [(oracle.bf_and_popcount(flt[3], f), oracle.bf_or_popcount(flt[3], f)) for f in flt]
    pairs = [(bfs[rng.randrange(n)], bfs[rng.randrange(n)]) for _ in range(17)]
    idx = {p: i for i, p in enumerate(bfs)}
    assert be.dist_pairs(pairs) == [(oracle.bf_and_popcount(flt[idx[a]], flt[idx[b]]), oracle.bf_or_popcount(flt[idx[a]], flt[idx[b]]))
                                    for a, b in pairs]
    want = np.array([[oracle.bf_and_popcount(a, b) for b in flt] for a in flt], dtype=np.int64)
    assert np.array_equal(be.dist_all_pairs(bfs), want)
    many = be.dist_all_pairs_many([bfs[:4], bfs[2:20], bfs[5:8]])
    assert np.array_equal(many[0], want[:4, :4]) and np.array_equal(many[1], want[2:20, 2:20]) and np.array_equal(many[2], want[5:8, 5:8])
    qb = be.query_positions_many(fas[:5])
    hits = be.fetch_hits([be.query_hits_many(qb, [4, 1], bfs)])[0]
    for r, q in enumerate([4, 1]):
        assert [int(x) for x in hits[r]] == [int(x) for x in oracle.query_hits(oracle.positions_distinct(open(fas[q], "rb").read(), k, num_bits), flt)]
    assert be.query_totals(qb) == [len(oracle.positions_distinct(open(f, "rb").read(), k, num_bits)) for f in fas[:5]]
    assert len(be._cache) <= 12


@pytest.mark.parametrize("key_bytes,digest_half,canonical", [("words", 0, "min"), ("quarter", 1, "min"), ("quarter", 0, "forward"), ("words", 1, "forward")])
def test_context_with_a_repinned_hash_spec(oracle, key_bytes, digest_half, canonical):
    """msbt_ctx_create(device, &spec): a context created with a non-default msbt_hash_spec builds the filters the oracle
    builds under the same switches (generic K1 kernel) -- re-pinning against a real howdesbt is configuration, not a rebuild."""
    from metasbt_b200 import ops
    rng = random.Random(17)
    spec = ops.HashSpec(key_bytes, digest_half, canonical)
    ctx2 = ops.Context(0, spec=spec)
    try:
        oracle.set_spec(key_bytes, digest_half, canonical)
        for k, num_bits in ((21, 1 << 16), (31, 200003), (33, 1 << 18), (60, 1 << 17), (64, 1 << 16)):
            fastas = [util.messy_fasta(rng, 3, 3000, "ACGTN") for _ in range(5)] + [b">e\n"]
            if k != 64:  # (at k = 64 the messy genomes hold no k-mer at all: the all-empty batch is a case of its own)
                fastas.append(util.codes_to_fasta(util.synthetic_codes(k, 9000), "clean"))
            for variant in (0, 3):
                _, got = build(ctx2, fastas, k, num_bits, variant=variant)
                for f, g in zip(fastas, got):
                    assert np.array_equal(g, oracle.makebf(f, k, num_bits))
        with pytest.raises(ValueError):
            build(ctx2, fastas, 21, 1 << 16, min_abund=2)
        got_spec = __import__("metasbt_b200._lib", fromlist=["x"]).HashSpecStruct()
        assert ctx2._L.msbt_ctx_get_spec(ctx2._h, ctypes.byref(got_spec)) == 0 and got_spec.digest_half == digest_half
    finally:
        oracle.set_spec()
        ctx2.close()


@pytest.mark.parametrize("na,nb,num_bits", [(1, 1, 64), (3, 17, 100000), (16, 16, 1 << 16), (17, 40, (1 << 18) + 64), (70, 5, 8192 * 3 + 64)])
def test_rectangular_all_pairs(ctx, oracle, na, nb, num_bits):
    """msbt_bf_and_popcount_rect: list A against list B (the block of shard.all_pairs_ring and of the tiled all-pairs path)."""
    rng = random.Random(na * 100 + nb)
    fastas = [util.messy_fasta(rng, 2, 2500, "ACGTN") for _ in range(na + nb)]
    mat, flt = build(ctx, fastas, 21, num_bits)
    a, b = [mat[i] for i in range(na)], [mat[na + j] for j in range(nb)]
    got = ctx.bf_rect(a, b, num_bits).cpu().numpy()
    want = np.array([[oracle.bf_and_popcount(flt[i], flt[na + j]) for j in range(nb)] for i in range(na)], dtype=np.int64)
    assert np.array_equal(got, want)
    assert ctx.bf_rect([], b, num_bits).shape == (0, nb)


@pytest.mark.parametrize("variant", [2, 3, 4])
@pytest.mark.parametrize("num_bits", [1 << 17, 200003, 1 << 20, (1 << 22) + 64])
def test_makebf_skewed_batch_after_a_random_one(ctx, oracle, variant, num_bits):
    """A skewed genome overflows its ring / scratch bucket and is completed by the direct redo pass, which can only ADD
    bits: whatever a rejected reservation leaves unwritten in the bucket must not be read as positions.  The scratch is
    first filled with the positions of a random batch of the same geometry, then low-complexity genomes are built."""
    rng = random.Random(num_bits + variant)
    noise = [util.codes_to_fasta(util.synthetic_codes(900 + i, 40_000), "n%d" % i) for i in range(6)]
    build(ctx, noise, 21, num_bits, variant=variant)
    skewed = [b">polyA\n" + b"A" * 30_000 + b"\n",
              b">two\n" + b"AC" * 20_000 + b"\n",
              util.messy_fasta(rng, 4, 6000, "ACGTN"),
              b">mix\n" + b"ACGT" * 3000 + b"N" + b"T" * 25_000 + b"\n"]
    for k in (1, 2, 5, 21):
        _, got = build(ctx, skewed, k, num_bits, variant=variant)
        for f, g in zip(skewed, got):
            assert np.array_equal(g, oracle.makebf(f, k, num_bits)), (k, f[:12])
        build(ctx, noise, 21, num_bits, variant=variant)


@pytest.mark.parametrize("mode", ["dense", "sparse", "hybrid", "auto"])
def test_sketch_pipeline_every_transport(ctx, oracle, mode):
    """pipeline.SketchPipeline: host FASTA bytes -> filters in host memory, chunked over three streams; the sink sees the
    oracle's filters whatever the transport (dense copy, position lists expanded on the host, both at once)."""
    from metasbt_b200 import pipeline
    rng = random.Random(len(mode))
    k, num_bits = 21, (1 << 20) + 64
    fastas = [util.messy_fasta(rng, 2, 6000, "ACGTN") for _ in range(23)] + [b">empty\n", util.codes_to_fasta(util.synthetic_codes(5, 30_000), "big")]
    offsets = [0]
    for f in fastas:
        offsets.append(offsets[-1] + len(f))
    text = torch.empty(offsets[-1], dtype=torch.uint8, pin_memory=True)
    text.numpy()[:] = np.frombuffer(b"".join(fastas), dtype=np.uint8)
    pipe = pipeline.SketchPipeline(ctx, k, num_bits, chunk_genomes=4, slots=2, mode=mode)
    got = {}
    seen_on_device = []

    def sink(first, words):
        for i in range(len(words)):
            got[first + i] = np.array(words[i], copy=True)

    for _ in range(2):  # the second pass reuses the slots (and, in hybrid mode, an adapted split)
        got.clear()
        pipe.run(text, offsets, sink=sink, on_device=lambda first, filt: seen_on_device.append((first, filt.shape[0])))
        assert sorted(got) == list(range(len(fastas)))
        for i, f in enumerate(fastas):
            assert np.array_equal(got[i], oracle.makebf(f, k, num_bits)), (mode, i)
    assert sum(n for _, n in seen_on_device) == 2 * len(fastas)
    assert pipe.h2d_bytes == 2 * offsets[-1] and pipe.d2h_bytes > 0
    if mode == "sparse":
        assert pipe.d2h_bytes < 2 * len(fastas) * (num_bits // 8)
    pipe.close()


@pytest.mark.parametrize("n_leaves", [2049, 5000, 17000])
@pytest.mark.parametrize("rows_kernel", ["1", "0"])
def test_query_sliced_rows_of_several_pieces(ctx, oracle, n_leaves, rows_kernel, monkeypatch):
    """More than 2,048 leaves: a row is several 256-byte pieces and the row-coherent probe (the warps of a CTA read the pieces
    of the same rows together) serves the query; MSBT_QS_ROWS=0 keeps the column-group kernel.  Same hits either way, equal
    to the oracle's; 17,000 leaves need more than the eight warps of one CTA."""
    from metasbt_b200 import ops
    monkeypatch.setenv("MSBT_QS_ROWS", rows_kernel)
    num_bits = 1 << 14
    rng = np.random.default_rng(n_leaves)
    w = ops.filter_words(num_bits)
    host = rng.integers(0, 1 << 63, size=(n_leaves, w), dtype=np.int64) & rng.integers(0, 1 << 63, size=(n_leaves, w), dtype=np.int64)
    dev = torch.from_numpy(host).to(ctx.device)
    leaves = [dev[i] for i in range(n_leaves)]
    leaf_np = [host[i].view(np.uint64) for i in range(n_leaves)]
    queries = [np.sort(rng.choice(num_bits, size=n, replace=False)).astype(np.uint64) for n in (9000, 4089, 0, 29, 4088)]
    positions = [torch.from_numpy(q.astype(np.uint32).view(np.int32)).to(ctx.device) for q in queries]
    want = np.stack([oracle.query_hits(q, leaf_np) for q in queries])
    sliced = ctx.bitslice(leaves, 0, num_bits)
    got = ctx.query_batch_sliced(sliced, n_leaves, 0, num_bits, positions).cpu().numpy()
    assert np.array_equal(got.astype(np.uint64), want)
    half = num_bits // 2
    part = ctx.query_batch_sliced(ctx.bitslice(leaves, 0, half), n_leaves, 0, half, positions) + \
        ctx.query_batch_sliced(ctx.bitslice(leaves, half, num_bits), n_leaves, half, num_bits, positions)
    assert np.array_equal(part.cpu().numpy().astype(np.uint64), want)


@pytest.mark.gpu
def test_allpairs_through_the_probe_equals_the_and_popcount_kernel(ctx, oracle, tmp_path, monkeypatch):
    """All pairs the K4 way (transpose once, every filter's set bits probe the bit-sliced matrix; work ~ set bits): the
    same integers as the K3 kernel -- for ragged sizes, several row passes, several query batches, an empty and a full
    filter -- and through the backend when the route is forced."""
    import torch
    from metasbt_b200 import backend as B, bffile, ops
    rng = np.random.default_rng(5)
    for num_bits, n, density in ((200_003, 70, 0.02), (1 << 20, 33, 0.05), (64 * 5, 9, 0.3)):
        words = ops.filter_words(num_bits)
        bits = rng.random((n, words * 64)) < density
        bits[:, num_bits:] = False
        bits[0, :] = False                      # an empty filter
        bits[1, :num_bits] = True               # a full one
        host = np.packbits(bits, axis=1, bitorder="little").view(np.uint64).reshape(n, words)
        dev = ctx.alloc_filters(n, num_bits)
        dev.zero_()
        dev[:, :words] = torch.from_numpy(host.view(np.int64)).to(dev.device)
        rows = [dev[i] for i in range(n)]
        want = ctx.bf_allpairs(rows, num_bits)
        row_bytes = ctx.sliced_row_words(n) * 4
        for slice_bytes, batch in ((48 << 30, 256), (row_bytes * (num_bits // 3 + 7), 16), (row_bytes * 64, 5)):
            got = ctx.bf_allpairs_probe(rows, num_bits, slice_bytes=slice_bytes, batch=batch)
            assert torch.equal(got, want), (num_bits, n, slice_bytes, batch)
        assert int(want[1, 1]) == num_bits and int(want[0].sum()) == 0
        assert np.array_equal(want.cpu().numpy()[2:6, 2:6],
                              np.array([[oracle.bf_and_popcount(host[i], host[j]) for j in range(2, 6)] for i in range(2, 6)]))
    # through the backend: forced routes agree, and the automatic rule stays with K3 for a handful of filters
    k, num_bits = 21, 1 << 16
    be = B.CudaBackend(k, num_bits)
    rnd = random.Random(9)
    fas, bfs = [], []
    for i in range(12):
        fa = tmp_path / ("p%02d.fa" % i)
        fa.write_bytes(util.messy_fasta(rnd, 2, 1500, "ACGTN"))
        fas.append(str(fa))
        bfs.append(str(tmp_path / ("p%02d.bf" % i)))
    be.sketch_many(fas, bfs, 1)
    monkeypatch.setenv("MSBT_ALLPAIRS_ROUTE", "bit")
    by_bit = be.dist_all_pairs(bfs)
    monkeypatch.setenv("MSBT_ALLPAIRS_ROUTE", "probe")
    before = be.ctx.launch_count
    by_probe = be.dist_all_pairs(bfs)
    assert be.ctx.launch_count - before >= 3     # transpose + extraction + probe
    monkeypatch.delenv("MSBT_ALLPAIRS_ROUTE")
    assert np.array_equal(by_bit, by_probe) and np.array_equal(be.dist_all_pairs(bfs), by_bit)
    flt = [bffile.read_bf(p)[1] for p in bfs]
    assert np.array_equal(by_probe, np.array([[oracle.bf_and_popcount(a, b) for b in flt] for a in flt], dtype=np.int64))
