"""The `metasbt update` path: Database.is_known / add(unknown) / characterize / _estimate_boundaries
/ update through the CLI, on the related-taxa data set of tests/dataset_update.py.

tests/golden/update_reference.json was produced by the UNMODIFIED reference metasbt/core.py driving a stand-in `howdesbt`
served by the oracle (tests/golden/make_reference_update_golden.py).  CPU: this repo's host logic over the oracle
backend must give the same database -- species assignments, new MSBT clusters, every .bf byte, clusters.tsv, genomes.tsv
and profile tables.  GPU: the shipped CUDA backend must give the same again."""
import json
import os

import pytest

from metasbt_b200 import cli, core
from tests import cpu_backend, dataset_update as D, snapshot
from tests.test_core import _diff, _order_insensitive

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "update_reference.json")


def _normalise(snap):
    """On top of test_core._order_insensitive: the reference writes `mags_list` by iterating a set difference
    (core.py:1556) and rebuilds `genomes` from children sets when a database is re-opened (core.py:231-246), so the order
    of MAGs inside clusters.tsv and of the rows of genomes.tsv depends on PYTHONHASHSEED; this code writes both sorted."""
    out = _order_insensitive(snap)
    rows = []
    for line in out["text"]["clusters.tsv"]:
        cols = line.split("\t")
        if len(cols) > 6 and not line.startswith("#"):
            cols[6] = ",".join(sorted(cols[6].split(",")))
        rows.append("\t".join(cols))
    out["text"]["clusters.tsv"] = rows
    head = [l for l in out["text"]["genomes.tsv"] if l.startswith("#")]
    out["text"]["genomes.tsv"] = head + sorted(l for l in out["text"]["genomes.tsv"] if not l.startswith("#"))
    return out


def _golden():
    with open(GOLDEN) as f:
        return json.load(f)


def test_update_matches_reference_python(tmp_path, oracle):
    snap = snapshot.run_update(str(tmp_path), backend=cpu_backend.OracleBackend(D.K, D.FILTER_BITS))
    gold = _golden()
    # every branch of the scenario is really exercised
    assert gold["update"]["species_assignments"]["MAG_a_known.fna"].endswith("s__Alphasynth_sp1")
    assert gold["update"]["unassigned"] == ["MAG_f_alien"]
    assert sorted(len(v) for v in gold["update"]["characterized"].values()) == [1, 3]
    assert snap["update"] == gold["update"]
    assert _diff(_normalise(snap), _normalise(gold)) == []


def test_estimate_boundaries_and_is_known(tmp_path, oracle):
    be = cpu_backend.OracleBackend(D.K, D.FILTER_BITS)
    refs_tsv, mags_txt, paths = D.write(str(tmp_path / "inputs"))
    cli.MetaSBT(["index", "--workdir", str(tmp_path), "--database", "db", "--references", refs_tsv, "--kmer-size", str(D.K),
                 "--filter-size", str(D.FILTER_BITS), "--min-kmer-occurrences", "1", "--nproc", "1"], backend=be)
    db = core.Database("db", str(tmp_path / "db"), str(tmp_path / "tmp"), backend=be)
    family = D.LINEAGE
    # species are fixed at 5 % (core.py:1320-1322)
    assert db._estimate_boundaries(family + "|g__Alphasynth|s__anything") == (0.0, 0.05)
    # above the species level the answer is the mean of the boundaries of the clusters of that level under the closest
    # ancestor that has any (core.py:1349-1386) -- for an existing genus too: taxonomy_exists() only accepts full
    # seven-level labels (core.py:348), so the report shortcut (core.py:1326-1334) is never taken for a partial label
    sib = [db.report[db.clusters["genus"][g].identifier]["boundaries"] for g in ("g__Alphasynth", "g__Betasynth")]
    mean = (round((sib[0][0] + sib[1][0]) / 2, 5), round((sib[0][1] + sib[1][1]) / 2, 5))
    assert db._estimate_boundaries(family + "|g__tmp") == mean
    assert db._estimate_boundaries(family + "|g__Alphasynth") == mean
    with pytest.raises(Exception, match="Unable to retrieve boundaries"):
        db._estimate_boundaries("k__Synthia")  # kingdoms: never estimated (core.py:1336-1337)
    with pytest.raises(Exception, match="Unable to retrieve boundaries"):
        db._estimate_boundaries("k__Nowhere")
    with pytest.raises(ValueError):
        db._estimate_boundaries("")
    # a reference genome is known by name; a MAG inside a species is assigned to it; a distant one is not
    ref0 = paths["SYN_000_ref"]
    assert db.is_known(ref0) == db.genomes["SYN_000_ref"].get_full_taxonomy()
    assert db.is_known(paths["MAG_a_known"]).endswith("|s__Alphasynth_sp1")
    assert db.is_known(paths["MAG_b_newsp"]) is None
    # add() of an unassigned genome keeps it for characterize(); update() refuses to run before that
    db.add(paths["MAG_b_newsp"], reference=False, taxonomy=None)
    db.add(paths["MAG_a_known"], reference=False, taxonomy=db.is_known(paths["MAG_a_known"]))
    with pytest.raises(Exception, match="unknown genomes in cache"):
        db.update()
    characterized, unassigned = db.characterize()
    assert unassigned == set() and list(characterized.values()) == [{"MAG_b_newsp"}]
    with pytest.raises(Exception, match="no unknown genomes"):
        db.characterize()
    db.update()
    assert "MAG_b_newsp" in db.genomes and not db.genomes["MAG_b_newsp"].is_known()


@pytest.mark.gpu
def test_update_cuda_backend_matches_reference_python(tmp_path, ctx):
    core._BACKENDS.clear()
    snap = snapshot.run_update(str(tmp_path))  # backend=None -> CudaBackend -> libmsbt.so
    gold = _golden()
    assert snap["update"] == gold["update"]
    assert _diff(_normalise(snap), _normalise(gold)) == []
