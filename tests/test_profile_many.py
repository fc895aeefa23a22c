"""Database.profile_many: the process pool of metasbt.py:1124-1143 (one `_profile` job per genome) as one batched,
level-synchronous descent.  CPU: equal to the per-genome method (which the reference-Python golden pins) with the oracle
backend, whatever the batch split.  GPU: >= 64 MAGs through `metasbt profile` with the CUDA backend against the same
database profiled with the oracle backend, table for table."""
import os
import shutil

import numpy as np
import pytest

from metasbt_b200 import cli, core
from tests import cpu_backend, dataset, util


def _extra_mags(folder: str, n: int):
    """n more MAGs: mutations (1-12 %) of the config-1 references, every eighth one unrelated."""
    refs = dataset.references()
    os.makedirs(folder, exist_ok=True)
    paths = []
    for m in range(n):
        name = "XMAG_%03d" % m
        if m % 8 == 7:
            codes = util.synthetic_codes(0xBEEF000 + m, 20_000 + 977 * m)
        else:
            src = (5 * m + 3) % len(refs)
            label, seed, length, _count = next(s for s in dataset.SPECIES if s[0] == refs[src][1])
            codes = util.mutate(util.synthetic_codes(seed, length), 0.01 + 0.11 * ((m * 37) % 100) / 100.0, 7000 + m)
            if m % 5 == 0:
                codes = codes[: len(codes) // 2]  # a fragmentary MAG
        p = os.path.join(folder, name + ".fna")
        with open(p, "wb") as f:
            f.write(dataset._contigs(codes, name, 1 + m % 4, m % 3 == 0))
        paths.append(p)
    return paths


def _index(workdir: str, backend):
    refs_tsv, _mags_txt, _ = dataset.write(os.path.join(workdir, "inputs"))
    app = cli.MetaSBT(["index", "--workdir", workdir, "--database", "db", "--references", refs_tsv,
                       "--kmer-size", str(dataset.K), "--filter-size", str(dataset.FILTER_BITS),
                       "--min-kmer-occurrences", str(dataset.MIN_OCC), "--nproc", "1"], backend=backend)
    return app.database


def _tables(workdir: str):
    d = os.path.join(workdir, "tmp", "profiles")
    return {fn: open(os.path.join(d, fn)).read() for fn in sorted(os.listdir(d))}


def test_profile_many_equals_per_genome_profile(tmp_path, oracle):
    be = cpu_backend.OracleBackend(dataset.K, dataset.FILTER_BITS)
    db = _index(str(tmp_path), be)
    mags = _extra_mags(str(tmp_path / "mags"), 20)
    entries = [core.Entry(db, os.path.splitext(os.path.basename(p))[0], os.path.splitext(os.path.basename(p))[0], "genome") for p in mags]
    sketches = core.Entry.sketch_many(entries, mags)
    batch = db.profile_many(mags, sketches, uncertainty=20.0)
    many = _tables(str(tmp_path))
    assert len(many) == 20
    # one by one, from scratch
    shutil.rmtree(tmp_path / "tmp" / "profiles")
    single = [db.profile(m, s, uncertainty=20.0) for m, s in zip(mags, sketches)]
    assert _tables(str(tmp_path)) == many
    assert single == batch
    # memoised tables are served without touching the backend (core.py:1932-1960)
    db._backend = None
    core._BACKENDS.clear()
    assert db.profile_many(mags, sketches, uncertainty=20.0) == batch
    db._backend = be
    # several batches (a tiny position budget) and another uncertainty / pruning threshold
    shutil.rmtree(tmp_path / "tmp" / "profiles")
    be.position_bytes = 4 * 250_000
    # A pruning threshold no phylum passes leaves nothing to descend into: the reference dies at the species level with the
    # IndexError of `sorted(...)[-1]` on an empty dictionary (core.py:2132).  Genome by genome first ...
    single, passing = dict(), []
    for i, (m, s) in enumerate(zip(mags, sketches)):
        try:
            single[i] = db.profile(m, s, uncertainty=50.0, pruning_threshold=0.3)
            passing.append(i)
        except IndexError:
            pass
    assert 5 < len(passing) < 20
    # ... then the batch: the same tables for the genomes that pass, the same exception as soon as one does not
    shutil.rmtree(tmp_path / "tmp" / "profiles")
    loose = db.profile_many([mags[i] for i in passing], [sketches[i] for i in passing], uncertainty=50.0, pruning_threshold=0.3)
    assert loose == [single[i] for i in passing]
    shutil.rmtree(tmp_path / "tmp" / "profiles")
    with pytest.raises(IndexError):
        db.profile_many(mags, sketches, uncertainty=50.0, pruning_threshold=0.3)
    del be.position_bytes
    with pytest.raises(ValueError):
        db.profile_many(mags, sketches[:-1])
    with pytest.raises(FileNotFoundError):  # no memoised table and no sketch (core.py:1700-1703)
        db.profile_many([mags[0]], [str(tmp_path / "nope.bf")])


@pytest.mark.gpu
def test_cli_profile_batch_of_72_mags_cuda_vs_oracle(tmp_path, ctx, oracle):
    from metasbt_b200 import ops
    core._BACKENDS.clear()
    want_dir, got_dir = str(tmp_path / "oracle"), str(tmp_path / "cuda")
    mags = _extra_mags(str(tmp_path / "mags"), 72)
    listing = str(tmp_path / "mags.txt")
    with open(listing, "w") as f:
        f.write("\n".join(mags) + "\n")
    _index(got_dir, None)  # CudaBackend -> libmsbt.so
    before = ops.default_context().launch_count
    got = cli.MetaSBT(["profile", "--workdir", got_dir, "--database", "db", "--genomes", listing, "--nproc", "1"])
    launches = ops.default_context().launch_count - before
    assert type(got.database.backend).__name__ == "CudaBackend"
    core._BACKENDS.clear()  # an injected backend registers itself for the static Database.dist: keep the two runs apart
    _index(want_dir, cpu_backend.OracleBackend(dataset.K, dataset.FILTER_BITS))
    want = cli.MetaSBT(["profile", "--workdir", want_dir, "--database", "db", "--genomes", listing, "--nproc", "1"],
                       backend=cpu_backend.OracleBackend(dataset.K, dataset.FILTER_BITS))
    assert len(got.result) == 72 and [t for _, t in got.result] == [t for _, t in want.result]
    assert _tables(got_dir) == _tables(want_dir)
    # batched: 72 genomes descend 8 tree levels with a few launches per level and tree, not a few per genome and tree
    # (the per-genome path needed > 20 launches per genome)
    assert 0 < launches < 72 * 6, launches
    # the sketches the batch wrote are the oracle's, byte for byte
    for p in mags[:8]:
        name = os.path.splitext(os.path.basename(p))[0] + ".bf"
        assert open(os.path.join(got_dir, "db", "sketches", name), "rb").read() == \
            open(os.path.join(want_dir, "db", "sketches", name), "rb").read()
