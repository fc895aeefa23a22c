"""The product on several GPUs (metasbt_b200/multigpu.py, MSBT_DEVICES): `metasbt index` / `profile` over two devices
reproduce the config-1 golden snapshot byte for byte; a large flat tree is probed through its bit-range-sharded sliced
matrix with an NCCL reduce and gives the oracle's hits.  Skipped on a one-GPU box."""
import json
import os
import random

import numpy as np
import pytest

from metasbt_b200 import core
from tests import snapshot, util

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _devices():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    return list(range(min(n, 4)))


def test_config1_on_several_gpus_matches_golden(tmp_path, ctx, monkeypatch):
    devs = _devices()
    monkeypatch.setenv("MSBT_DEVICES", ",".join(str(d) for d in devs))
    core._BACKENDS.clear()
    try:
        snap = snapshot.run_config1(str(tmp_path))
        be = core._BACKENDS[(31, 1 << 22)]
        assert type(be).__name__ == "MultiGpuBackend" and be.P == len(devs)
        # the work really was spread: every device built sketches and launched kernels
        assert all(p.ctx.launch_count > 0 for p in be.parts)
        assert len(set(be._home.values())) == len(devs)
        with open(os.path.join(GOLDEN, "config1.json")) as f:
            assert snap == json.load(f)
    finally:
        core._BACKENDS.clear()


def test_update_path_on_several_gpus_matches_golden(tmp_path, ctx, monkeypatch):
    devs = _devices()
    monkeypatch.setenv("MSBT_DEVICES", ",".join(str(d) for d in devs))
    core._BACKENDS.clear()
    try:
        snap = snapshot.run_update(str(tmp_path))
        with open(os.path.join(GOLDEN, "update_reference.json")) as f:
            want = json.load(f)
        for key in ("filters", "profiles", "update", "metadata_raw"):
            assert snap[key] == want[key], key
    finally:
        core._BACKENDS.clear()


def test_large_tree_is_sharded_by_bit_range_and_reduced_with_nccl(tmp_path, ctx, oracle):
    from metasbt_b200 import bffile, multigpu
    devs = _devices()
    rng = random.Random(77)
    k, num_bits, n_leaves = 21, 1 << 17, 300
    be = multigpu.MultiGpuBackend(k, num_bits, devices=devs)
    fas, bfs = [], []
    for i in range(n_leaves):
        fa = tmp_path / ("leaf%03d.fa" % i)
        fa.write_bytes(util.messy_fasta(rng, 1, 1500, "ACGT"))
        fas.append(str(fa))
        bfs.append(str(tmp_path / ("leaf%03d.bf" % i)))
    be.sketch_many(fas, bfs, 1)
    leaf_np = [bffile.read_bf(p)[1] for p in bfs]
    for f, fa in zip(leaf_np[:20], fas[:20]):
        assert np.array_equal(f, oracle.makebf(open(fa, "rb").read(), k, num_bits))
    queries = []
    for q in range(9):
        qfa = tmp_path / ("q%d.fa" % q)
        qfa.write_bytes(open(fas[(31 * q) % n_leaves], "rb").read() + util.messy_fasta(rng, 2, 700, "ACGTN"))
        queries.append(str(qfa))
    qb = be.query_positions_many(queries)
    want_pos = [oracle.positions_distinct(open(q, "rb").read(), k, num_bits) for q in queries]
    assert be.query_totals(qb) == [len(p) for p in want_pos]
    assert len(set(qb.where)) == len(devs)  # the queries were hashed on every device
    idxs = [8, 0, 3, 4, 7]
    hits = be.fetch_hits([be.query_hits_many(qb, idxs, bfs), be.query_hits_many(qb, [1, 2], bfs[:40])])
    assert be.sliced_probes == len(idxs) and be.nccl_reduces == 1 and len(be._trees) == 1
    want = np.stack([oracle.query_hits(want_pos[i], leaf_np) for i in idxs])
    assert np.array_equal(hits[0].astype(np.uint64), want)
    assert np.array_equal(hits[1].astype(np.uint64), np.stack([oracle.query_hits(want_pos[i], leaf_np[:40]) for i in (1, 2)]))
    tree = next(iter(be._trees.values()))
    assert [t.device.index for t in tree.sliced] == devs and sum(t.shape[0] for t in tree.sliced) == num_bits
    # K2 / K3 across devices
    be.union(bfs, str(tmp_path / "u.bf"))
    assert np.array_equal(bffile.read_bf(str(tmp_path / "u.bf"))[1], oracle.bf_or(leaf_np))
    assert be.popcounts(bfs[:50]) == [oracle.bf_popcount(f) for f in leaf_np[:50]]
    assert be.dist_one_vs_many(bfs[3], bfs[:30]) == [(oracle.bf_and_popcount(leaf_np[3], f), oracle.bf_or_popcount(leaf_np[3], f)) for f in leaf_np[:30]]
    got = be.dist_all_pairs(bfs[10:22])
    assert np.array_equal(got, np.array([[oracle.bf_and_popcount(a, b) for b in leaf_np[10:22]] for a in leaf_np[10:22]], dtype=np.int64))
    # rewriting a leaf drops the sharded matrix and every replica
    be.sketch_many([fas[7]], [bfs[5]], 1)
    assert len(be._trees) == 0
    hits2 = be.fetch_hits([be.query_hits_many(qb, [0], bfs)])[0]
    assert int(hits2[0][5]) == int(hits2[0][7])
