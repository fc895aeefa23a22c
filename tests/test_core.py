"""Config 1 (BASELINE.json configs[0] stand-in): `metasbt index` + `metasbt profile` through the CLI and
the Database/Entry API.  CPU: host logic served by the oracle backend must reproduce the committed golden
snapshot.  GPU: the shipped CUDA backend must reproduce the same snapshot byte for byte."""
import json
import os

import numpy as np
import pytest

from metasbt_b200 import bffile, cli, core
from tests import cpu_backend, dataset, snapshot

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _golden():
    with open(os.path.join(GOLDEN, "config1.json")) as f:
        return json.load(f)


def _diff(a, b, path=""):
    if type(a) != type(b):
        return ["%s: type %s != %s" % (path, type(a).__name__, type(b).__name__)]
    if isinstance(a, dict):
        out = []
        for k in sorted(set(a) | set(b)):
            if k not in a or k not in b:
                out.append("%s/%s: only on one side" % (path, k))
            else:
                out.extend(_diff(a[k], b[k], path + "/" + k))
        return out
    if a != b:
        return ["%s: %r != %r" % (path, a, b)]
    return []


def _reference_golden():
    with open(os.path.join(GOLDEN, "config1_reference.json")) as f:
        return json.load(f)


def _order_insensitive(snap):
    """The reference writes child lists by iterating a Python set (core.py:3746-3776, 3870-3875: order depends on
    PYTHONHASHSEED); this code writes them sorted.  Compare `<name>.txt` and `index.detbrief.sbt` as: first line of the
    tree (the cluster's own filter) + the multiset of the remaining lines."""
    out = json.loads(json.dumps(snap))
    for rel, lines in out["text"].items():
        if rel.endswith(".sbt"):
            out["text"][rel] = [lines[0]] + sorted(lines[1:])
        elif rel.endswith(".txt") and rel.startswith("clusters"):
            out["text"][rel] = sorted(lines)
    return out


def test_config1_matches_reference_python(tmp_path, oracle):
    """Tier-2 parity: tests/golden/config1_reference.json was produced by the UNMODIFIED reference
    metasbt/core.py (Database.set_configs/add/update/_profile, Entry.sketch/index/get_density/get_boundaries, Database.dist)
    driving a stand-in `howdesbt` served by the oracle (tests/golden/make_reference_golden.py).  This repo's own
    Database/Entry/CLI must give the same database: every .bf byte, clusters.tsv (densities, centroids, ANI boundaries),
    genomes.tsv, metadata.json and every profile table -- identical up to the order of child lines."""
    snap = snapshot.run_config1(str(tmp_path), backend=cpu_backend.OracleBackend(dataset.K, dataset.FILTER_BITS))
    assert _diff(_order_insensitive(snap), _order_insensitive(_reference_golden())) == []


def test_config1_oracle_backend_matches_golden(tmp_path, oracle):
    snap = snapshot.run_config1(str(tmp_path), backend=cpu_backend.OracleBackend(dataset.K, dataset.FILTER_BITS))
    assert _diff(snap, _golden()) == []


def test_config1_layout(tmp_path, oracle):
    """On-disk contract of SURVEY.md App. C."""
    snapshot.run_config1(str(tmp_path), backend=cpu_backend.OracleBackend(dataset.K, dataset.FILTER_BITS))
    db = tmp_path / "db"
    # metadata.json: key insertion order of core.py:153-155, 420, 470, 530 and plain json.dump
    assert (db / "metadata.json").read_text() == json.dumps(
        {"clusters_count": 15, "flat": True, "min_kmer_occurrence": 1, "kmer_size": 31, "filter_size": 1 << 22})
    assert len(list((db / "sketches").glob("*.bf"))) == 12 + 6
    for bf in (db / "sketches").glob("*.bf"):
        assert bf.stat().st_size == 112 + 8 + (1 << 22) // 8
    # MSBT1 is the first species created; cluster folders are clusters/<batch6>/<MSBTn>/
    sp = db / "clusters" / "000000" / "MSBT1"
    assert (sp / "s__Monkeypox_virus.bf").is_file() and (sp / "s__Monkeypox_virus.txt").is_file()
    tree = (sp / "tree" / "index.detbrief.sbt").read_text().split("\n")
    assert tree[0] == str(sp / "s__Monkeypox_virus.bf") and all(t.startswith("*" + str(db / "sketches")) for t in tree[1:5])
    assert (db / "clusters" / "000000" / "MSBT0" / "db.bf").is_file()
    # a species with < 3 genomes has no boundaries (core.py:4019-4024)
    report = core.Database._load_report(str(db / "clusters.tsv"))
    assert report["MSBT8"]["boundaries"] == (None, None) and report["MSBT1"]["boundaries"][0] is not None
    # the database re-opens from disk and profiles from the memoised table (core.py:1932-1954)
    again = core.Database("db", str(db), str(tmp_path / "tmp"), backend=cpu_backend.OracleBackend(dataset.K, dataset.FILTER_BITS))
    assert again.size() == (1, 2, 2, 2, 2, 2, 4) and len(again.genomes) == 12
    mag = str(tmp_path / "inputs" / "genomes" / "GCA_011395005.1_mag0.fna")
    prof = again.profile(mag, str(db / "sketches" / "GCA_011395005.1_mag0.bf"))
    assert list(prof["genome"].values()) == [0.01933]


def test_cluster_filter_is_union_and_copy(tmp_path, oracle):
    snapshot.run_config1(str(tmp_path), backend=cpu_backend.OracleBackend(dataset.K, dataset.FILTER_BITS))
    db = tmp_path / "db"
    sk = sorted((db / "sketches").glob("*ref[0-3].bf"))
    assert len(sk) == 4
    union = oracle.bf_or([bffile.read_bf(str(p))[1] for p in sk])
    assert np.array_equal(bffile.read_bf(str(db / "clusters/000000/MSBT1/s__Monkeypox_virus.bf"))[1], union)
    # single-child cluster = byte copy of the child (core.py:3781-3789)
    a = (db / "clusters/000000/MSBT11/g__Orthopneumovirus.bf").read_bytes()
    assert a == (db / "clusters/000000/MSBT10/s__Human_orthopneumovirus.bf").read_bytes()


def test_cli_rejects_what_is_out_of_scope(tmp_path):
    refs, _mags, _ = dataset.write(str(tmp_path / "in"))
    base = ["index", "--workdir", str(tmp_path / "w"), "--database", "db", "--references", refs, "--kmer-size", "31",
            "--filter-size", "4096"]
    with pytest.raises(NotImplementedError):
        cli.MetaSBT(base + ["--pack"], backend=object())
    with pytest.raises(FileNotFoundError):
        cli.MetaSBT(["profile", "--workdir", str(tmp_path / "nope"), "--database", "db", "--genome", refs])


def test_no_cpu_fallback_in_product_path():
    """The shipped package must not import the oracle or carry a CPU implementation."""
    import pathlib
    pkg = pathlib.Path(core.__file__).parent
    for p in pkg.glob("*.py"):
        text = p.read_text()
        assert "from oracle" not in text and "import oracle" not in text, p


@pytest.mark.gpu
def test_config1_cuda_backend_matches_golden(tmp_path, ctx):
    core._BACKENDS.clear()
    snap = snapshot.run_config1(str(tmp_path))  # backend=None -> CudaBackend -> libmsbt.so
    from metasbt_b200 import ops
    assert ops.default_context().launch_count > 0  # the CUDA kernels really ran
    assert _diff(snap, _golden()) == []
    assert _diff(_order_insensitive(snap), _order_insensitive(_reference_golden())) == []  # == the reference's own Python
