"""`metasbt index` / `profile` at the reference CLI's DEFAULT `--min-kmer-occurrences 2` (metasbt.py:416-425): only k-mers seen
at least twice in a genome enter its filter (`howdesbt makebf --min=2`, core.py:3920-3931).  The config-1 stand-in gets a
second record per genome that repeats 45 % of its sequence, so that the filters are not empty.  CPU: the database builds
over the oracle backend and differs from the `--min 1` one; GPU: the CUDA backend (exact-multiplicity K1 path: bit table,
candidate table, shared-memory candidate bitmap, three genomes in flight) reproduces it byte for byte."""
import os

import pytest

from metasbt_b200 import cli
from tests import cpu_backend, dataset, snapshot


def _write_with_repeats(folder):
    refs_tsv, mags_txt, paths = dataset.write(folder)
    for name, p in paths.items():
        with open(p, "rb") as f:
            lines = f.read().split(b"\n")
        seq = [ln for ln in lines if ln and not ln.startswith(b">")]
        with open(p, "ab") as f:
            f.write(b">" + name.encode() + b"_repeat\n" + b"\n".join(seq[: int(len(seq) * 0.45)]) + b"\n")
    return refs_tsv, mags_txt


def _run(workdir, backend, min_occ):
    refs_tsv, mags_txt = _write_with_repeats(os.path.join(workdir, "inputs"))
    cli.MetaSBT(["index", "--workdir", workdir, "--database", "db", "--references", refs_tsv, "--kmer-size", str(dataset.K),
                 "--filter-size", str(dataset.FILTER_BITS), "--min-kmer-occurrences", str(min_occ), "--nproc", "1"], backend=backend)
    prof = cli.MetaSBT(["profile", "--workdir", workdir, "--database", "db", "--genomes", mags_txt, "--nproc", "1"], backend=backend)
    assert len(prof.result) == 6
    return snapshot.snapshot(workdir)


def _densities(snap):
    rows = [ln.split("\t") for ln in snap["text"]["clusters.tsv"] if ln and not ln.startswith("#")]
    header = [ln for ln in snap["text"]["clusters.tsv"] if ln.startswith("# cluster_id")][0][2:].split("\t")
    col = header.index("density")
    return [float(r[col]) for r in rows]


def test_min2_database_over_the_oracle(tmp_path, oracle):
    two = _run(str(tmp_path / "two"), cpu_backend.OracleBackend(dataset.K, dataset.FILTER_BITS), 2)
    one = _run(str(tmp_path / "one"), cpu_backend.OracleBackend(dataset.K, dataset.FILTER_BITS), 1)
    assert '"min_kmer_occurrence": 2' in two["metadata_raw"]
    d2, d1 = _densities(two), _densities(one)
    assert len(d2) == len(d1) == 15 and all(x > 0 for x in d2)
    assert all(a < b for a, b in zip(sorted(d2), sorted(d1)))       # about 45 % of the k-mers survive the threshold
    assert two["filters"].keys() == one["filters"].keys()
    assert all(two["filters"][k]["sha256_payload"] != one["filters"][k]["sha256_payload"] for k in two["filters"])


@pytest.mark.gpu
def test_min2_database_cuda_equals_oracle(tmp_path, oracle):
    from metasbt_b200 import backend as cuda_backend, core
    core._BACKENDS.clear()
    cuda = cuda_backend.CudaBackend(dataset.K, dataset.FILTER_BITS)
    before = cuda.ctx.launch_count
    got = _run(str(tmp_path / "cuda"), cuda, 2)
    assert cuda.ctx.launch_count > before
    core._BACKENDS.clear()
    want = _run(str(tmp_path / "oracle"), cpu_backend.OracleBackend(dataset.K, dataset.FILTER_BITS), 2)
    assert got == want
