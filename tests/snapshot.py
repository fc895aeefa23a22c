"""TEST INFRASTRUCTURE: run `metasbt index` + `metasbt profile` on the config-1 stand-in with a given
backend and reduce the resulting database to a JSON-able snapshot (path- and date-independent) that can
be compared across backends (oracle vs CUDA) and against tests/golden/config1.json."""
import hashlib
import json
import os
from typing import Any, Dict

from metasbt_b200 import bffile, cli
from tests import dataset, dataset_update


def _rel(text: str, root: str) -> str:
    return text.replace(root, "<W>")


def run_config1(workdir: str, backend=None) -> Dict[str, Any]:
    refs_tsv, mags_txt, _ = dataset.write(os.path.join(workdir, "inputs"))
    argv = ["index", "--workdir", workdir, "--database", "db", "--references", refs_tsv,
            "--kmer-size", str(dataset.K), "--filter-size", str(dataset.FILTER_BITS),
            "--min-kmer-occurrences", str(dataset.MIN_OCC), "--nproc", "1"]
    app = cli.MetaSBT(argv, backend=backend)
    prof = cli.MetaSBT(["profile", "--workdir", workdir, "--database", "db", "--genomes", mags_txt, "--nproc", "1"],
                       backend=backend)
    assert app.database is not None and len(prof.result) == 6
    return snapshot(workdir)


def run_update(workdir: str, backend=None) -> Dict[str, Any]:
    """`metasbt index` on the related-taxa data set, then `metasbt update` with its six MAGs."""
    refs_tsv, mags_txt, _ = dataset_update.write(os.path.join(workdir, "inputs"))
    cli.MetaSBT(["index", "--workdir", workdir, "--database", "db", "--references", refs_tsv,
                 "--kmer-size", str(dataset_update.K), "--filter-size", str(dataset_update.FILTER_BITS),
                 "--min-kmer-occurrences", str(dataset_update.MIN_OCC), "--nproc", "1"], backend=backend)
    app = cli.MetaSBT(["update", "--workdir", workdir, "--database", "db", "--genomes", mags_txt, "--nproc", "1"],
                      backend=backend)
    species_assignments, characterized, unassigned = app.result
    snap = snapshot(workdir)
    snap["update"] = {
        "species_assignments": {os.path.basename(k): v for k, v in species_assignments.items()},
        "characterized": {k: sorted(v) for k, v in characterized.items()},
        "unassigned": sorted(unassigned)}
    return json.loads(json.dumps(snap, sort_keys=True))


def snapshot(workdir: str) -> Dict[str, Any]:
    db = os.path.join(workdir, "db")
    snap: Dict[str, Any] = {"filters": {}, "text": {}, "profiles": {}}
    with open(os.path.join(db, "metadata.json")) as f:
        snap["metadata_raw"] = f.read()
    for name in ("clusters.tsv", "genomes.tsv"):
        with open(os.path.join(db, name)) as f:
            lines = f.read().split("\n")
        assert lines[0].startswith("# MetaSBT ")
        snap["text"][name] = lines[1:]  # the first line carries today's date (core.py:1522)
    for root, _dirs, files in os.walk(db):
        for fn in sorted(files):
            p = os.path.join(root, fn)
            rel = os.path.relpath(p, db)
            if fn.endswith(".bf"):
                header, words = bffile.read_bf(p)
                snap["filters"][rel] = {
                    "sha256_payload": hashlib.sha256(words.tobytes()).hexdigest(),
                    "sha256_file": hashlib.sha256(open(p, "rb").read()).hexdigest(),
                    "k": header.kmer_size, "bits": header.num_bits, "size": os.path.getsize(p)}
            elif fn.endswith(".txt") or fn.endswith(".sbt"):
                with open(p) as f:
                    snap["text"][rel] = _rel(f.read(), workdir).split("\n")
    prof_dir = os.path.join(workdir, "tmp", "profiles")
    for fn in sorted(os.listdir(prof_dir)):
        with open(os.path.join(prof_dir, fn)) as f:
            snap["profiles"][fn] = f.read().split("\n")
    return json.loads(json.dumps(snap, sort_keys=True))
