"""flat=False trees: `howdesbt cluster` as greedy Hamming agglomeration over the backend's batch
operators (metasbt_b200/cluster.py), against the oracle's brute-force twin; a clustered database answers `profile` exactly
like the flat one; `dereplicate` (core.py:2535-2741).  CPU with the oracle backend, GPU with the CUDA backend."""
import os
import random

import numpy as np
import pytest

from metasbt_b200 import bffile, cli, cluster, core
from tests import cpu_backend, dataset, util


def _leaf_files(folder, n, k, num_bits, rng, backend):
    os.makedirs(folder, exist_ok=True)
    groups = [util.synthetic_codes(500 + g, 6000) for g in range(max(2, n // 3))]
    fas, bfs = [], []
    for i in range(n):
        codes = util.mutate(groups[i % len(groups)], 0.002 * (1 + i % 5), 40 + i)
        fa = os.path.join(folder, "leaf%02d.fa" % i)
        with open(fa, "wb") as f:
            f.write(util.codes_to_fasta(codes, "leaf%02d" % i))
        fas.append(fa)
        bfs.append(os.path.join(folder, "leaf%02d.bf" % i))
    backend.sketch_many(fas, bfs, 1)
    return bfs


def _check_topology(folder, backend, oracle, n, k=21, num_bits=1 << 15):
    bfs = _leaf_files(folder, n, k, num_bits, random.Random(n), backend)
    topo = cluster.cluster_tree(backend, bfs, os.path.join(folder, "node{number}"))
    want = oracle.cluster_topology([bffile.read_bf(p)[1] for p in bfs])
    assert [d for d, _ in topo] == [d for d, _ in want]
    number = 0
    for (depth, path), (_d, node) in zip(topo, want):
        if node < n:
            assert path == bfs[node]
        else:
            number += 1
            assert path == os.path.join(folder, "node%d.bf" % number) and os.path.isfile(path)
    # every internal node is the union of the leaves below it
    for i, (depth, path) in enumerate(topo):
        below = [p for d, p in _subtree(topo, i) if p in bfs]
        if path not in bfs:
            assert np.array_equal(bffile.read_bf(path)[1], oracle.bf_or([bffile.read_bf(p)[1] for p in below]))
    assert cluster.topology_leaves(topo) == [p for _, p in topo if p in bfs] and sorted(cluster.topology_leaves(topo)) == sorted(bfs)
    assert not [f for f in os.listdir(folder) if "tmp" in f]
    cluster.write_topology(os.path.join(folder, "t.sbt"), topo)
    assert cluster.read_topology(os.path.join(folder, "t.sbt")) == topo


def _subtree(topo, i):
    out = [topo[i]]
    for d, p in topo[i + 1:]:
        if d <= topo[i][0]:
            break
        out.append((d, p))
    return out


def test_greedy_merges_and_preorder_small_cases():
    # three leaves, 0 and 2 closest; then (1, 3)
    ham = [[0, 9, 2], [9, 0, 7], [2, 7, 0]]
    made = []
    merges = cluster.greedy_merges(ham, lambda u, v, w: made.append((u, v, w)), lambda w, others: [5 for _ in others])
    assert merges == [(0, 2), (1, 3)] and made == [(0, 2, 3), (1, 3, 4)]
    assert cluster.preorder(3, merges) == [(0, 4), (1, 1), (1, 3), (2, 0), (2, 2)]
    assert cluster.preorder(1, []) == [(0, 0)] and cluster.preorder(0, []) == []
    # ties: the smaller node ids win
    ham = [[0, 4, 4], [4, 0, 4], [4, 4, 0]]
    assert cluster.greedy_merges(ham, lambda *a: None, lambda w, o: [1 for _ in o])[0] == (0, 1)
    with pytest.raises(ValueError):
        cluster.cluster_tree(None, ["a", "b"], "no_placeholder")


@pytest.mark.parametrize("n", [1, 2, 3, 7, 12])
def test_cluster_tree_oracle_backend_vs_bruteforce_twin(tmp_path, oracle, n):
    _check_topology(str(tmp_path), cpu_backend.OracleBackend(21, 1 << 15), oracle, n)


@pytest.mark.gpu
@pytest.mark.parametrize("n", [2, 9, 33])
def test_cluster_tree_cuda_backend_vs_bruteforce_twin(tmp_path, ctx, oracle, n):
    from metasbt_b200 import backend as B
    _check_topology(str(tmp_path), B.CudaBackend(21, 1 << 15), oracle, n)


def _profiles(workdir, flat, backend):
    refs_tsv, mags_txt, paths = dataset.write(os.path.join(workdir, "inputs"))
    db = core.Database("db", os.path.join(workdir, "db"), os.path.join(workdir, "tmp"), flat=flat, nproc=1, backend=backend)
    refs = dict(line.strip().split("\t") for line in open(refs_tsv) if line.strip() and not line.startswith("#"))
    db.set_configs(sorted(refs), min_kmer_occurrence=1, kmer_size=dataset.K, filter_size=dataset.FILTER_BITS)
    db.add_many([(g, True, refs[g]) for g in sorted(refs)])
    db.update()
    mags = [line.strip() for line in open(mags_txt) if line.strip()]
    entries = [core.Entry(db, os.path.splitext(os.path.basename(p))[0], os.path.splitext(os.path.basename(p))[0], "genome") for p in mags]
    tables = db.profile_many(mags, core.Entry.sketch_many(entries, mags), uncertainty=20.0)
    return db, tables


def _check_clustered_database(tmp_path, backend_factory):
    flat_db, flat_tables = _profiles(str(tmp_path / "flat"), True, backend_factory())
    core._BACKENDS.clear()
    deep_db, deep_tables = _profiles(str(tmp_path / "deep"), False, backend_factory())
    assert deep_tables == flat_tables  # union nodes only prune: the leaves, and so every hit count, are the same
    assert '"flat": false' in open(os.path.join(deep_db.root, "metadata.json")).read()
    # the Monkeypox species (4 genomes) has a 3-level binary tree: 3 internal nodes + 4 leaves
    sp = deep_db.clusters["species"]["s__Monkeypox_virus"]
    topo = cluster.read_topology(os.path.join(sp.folder, "tree", "index.detbrief.sbt"))
    assert len(topo) == 7 and topo[0] == (0, os.path.join(sp.folder, "tree", "node1.bf"))
    assert sorted(os.path.basename(p) for p in cluster.topology_leaves(topo)) == sorted(g + ".bf" for g in sp.children)
    assert not os.path.exists(os.path.join(sp.folder, "tree", "union.sbt"))
    # the same reports otherwise (densities, centroids, boundaries do not depend on the tree shape)
    assert open(os.path.join(deep_db.root, "clusters.tsv")).read().split("\n")[1:] == \
        open(os.path.join(flat_db.root, "clusters.tsv")).read().split("\n")[1:]
    # a clustered database re-opens from disk
    again = core.Database("db", deep_db.root, deep_db.tmp, backend=backend_factory())
    assert again.flat is False and again.size() == deep_db.size()


def test_clustered_database_profiles_like_the_flat_one(tmp_path, oracle):
    _check_clustered_database(tmp_path, lambda: cpu_backend.OracleBackend(dataset.K, dataset.FILTER_BITS))


@pytest.mark.gpu
def test_clustered_database_cuda(tmp_path, ctx, oracle):
    from metasbt_b200 import backend as B
    _check_clustered_database(tmp_path, lambda: B.CudaBackend(dataset.K, dataset.FILTER_BITS))


def _check_dereplicate(tmp_path, backend):
    refs_tsv, _mags_txt, paths = dataset.write(str(tmp_path / "inputs"))
    extra = str(tmp_path / "inputs" / "genomes")
    names = sorted(paths)
    ref0 = paths["GCA_024732825.1_ref0"]
    # an exact copy of ref0 (distance 0) and a 0.05 % mutant of it (distance ~ 0.0005) under names that sort AFTER ref0
    twin, near = os.path.join(extra, "ZZ_twin.fna"), os.path.join(extra, "ZZ_near.fna")
    open(twin, "wb").write(open(ref0, "rb").read())
    label, seed, length, _count = dataset.SPECIES[0]
    codes = util.mutate(util.mutate(util.synthetic_codes(seed, length), 0.01, 1000), 0.0005, 31337)
    open(near, "wb").write(util.codes_to_fasta(codes, "near"))
    db = core.Database("db", str(tmp_path / "db"), str(tmp_path / "tmp"), flat=True, nproc=1, backend=backend)
    refs = dict(line.strip().split("\t") for line in open(refs_tsv) if line.strip() and not line.startswith("#"))
    db.set_configs(sorted(refs), min_kmer_occurrence=1, kmer_size=dataset.K, filter_size=dataset.FILTER_BITS)
    genomes = set(refs) | {twin, near}
    kept = db.dereplicate(genomes, threshold=0.002)
    assert kept == set(refs)                      # both replicas of ref0 go, every reference stays (they differ by >= 1 %)
    assert db.dereplicate(genomes, threshold=0.0001) == genomes - {twin}
    assert db.dereplicate({ref0}, threshold=0.5) == {ref0}
    with pytest.raises(ValueError):
        db.dereplicate(genomes, threshold=1.0)
    with pytest.raises(ValueError):
        db.dereplicate(genomes, compare_with="nothing")
    with pytest.raises(Exception):
        db.dereplicate(set())
    # against the database: index the references, then the twin and the near copy are replicas, a MAG 2 % away is not
    db.add_many([(g, True, refs[g]) for g in sorted(refs)])
    db.update()
    mag0 = paths["GCA_011395005.1_mag0"]
    assert db.dereplicate({twin, near, mag0}, threshold=0.002, compare_with="database") == {mag0}
    # through the CLI: --dereplicate on index drops the copies before they are added
    with open(refs_tsv, "a") as f:
        f.write("%s\t%s\n%s\t%s\n" % (twin, label, near, label))
    app = cli.MetaSBT(["index", "--workdir", str(tmp_path / "w2"), "--database", "db", "--references", refs_tsv, "--kmer-size",
                       str(dataset.K), "--filter-size", str(dataset.FILTER_BITS), "--min-kmer-occurrences", "1", "--dereplicate", "0.002"],
                      backend=backend)
    assert len(app.database.genomes) == 12 and "ZZ_twin" not in app.database.genomes


def test_dereplicate_oracle_backend(tmp_path, oracle):
    _check_dereplicate(tmp_path, cpu_backend.OracleBackend(dataset.K, dataset.FILTER_BITS))


@pytest.mark.gpu
def test_dereplicate_cuda_backend(tmp_path, ctx, oracle):
    from metasbt_b200 import backend as B
    core._BACKENDS.clear()
    _check_dereplicate(tmp_path, B.CudaBackend(dataset.K, dataset.FILTER_BITS))
