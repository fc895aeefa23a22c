"""Filter-size estimation (reference core.py:470-533 runs ntCard): the distinct k-mer count comes from
K1 with modulus sampling + K3' + linear counting.  CPU: the estimator's arithmetic and Database.set_configs over the
oracle counter, against an exact count.  GPU: the CUDA counter returns the very popcount the oracle counter returns."""
import json
import math
import os

import pytest

from metasbt_b200 import core, estimate
from tests import cpu_backend, dataset


def _exact_distinct(paths, k, oracle):
    seen = set()
    for p in paths:
        for w in oracle.py_kmers(open(p, "rb").read(), k):
            seen.add(oracle.py_canonical(w))
    return len(seen)


def test_linear_count_and_sampling_factor():
    assert estimate.linear_count(0, 1 << 20) == 0.0
    m = 1 << 20
    for n in (1000, 50_000, 200_000):
        x = round(m * (1.0 - math.exp(-n / m)))  # expected set bits for n distinct items
        assert abs(estimate.linear_count(x, m) - n) / n < 1e-3
    with pytest.raises(ValueError):
        estimate.linear_count(m, m)
    with pytest.raises(ValueError):
        estimate.linear_count(m + 1, m)
    assert estimate.sampling_factor(1000) == 1
    assert estimate.sampling_factor(10 * estimate.EST_BITS) == 80
    assert estimate.filter_size_from_f0(1000, 50.0) == 1500 and estimate.filter_size_from_f0(1000, None) == 1000
    with pytest.raises(ValueError):
        estimate.filter_size_from_f0(0, 50.0)


@pytest.mark.parametrize("num_bits,forced_s", [(1 << 24, None), (1 << 20, 4)])
def test_estimate_against_exact_count(tmp_path, oracle, num_bits, forced_s, monkeypatch):
    _, _, _ = dataset.write(str(tmp_path / "inputs"))
    paths = sorted(str(p) for p in (tmp_path / "inputs").rglob("*.fna"))[:6]
    assert paths
    exact = _exact_distinct(paths, dataset.K, oracle)
    if forced_s:
        monkeypatch.setattr(estimate, "sampling_factor", lambda total, bits=estimate.EST_BITS: forced_s)
    backend = cpu_backend.OracleBackend(dataset.K, dataset.FILTER_BITS)
    f0, s = estimate.estimate_distinct_kmers(paths, dataset.K, backend.count_union_bits, num_bits=num_bits)
    assert s == (forced_s or 1)
    # linear counting at load <= 1/8 plus (for S > 1) binomial sampling noise ~ sqrt(S / n)
    tol = 0.005 if not forced_s else 0.02
    assert abs(f0 - exact) / exact < tol, (f0, exact)


def test_set_configs_estimates_the_filter_size(tmp_path, oracle, monkeypatch):
    refs_tsv, _, _ = dataset.write(str(tmp_path / "inputs"))
    paths = sorted(str(p) for p in (tmp_path / "inputs").rglob("*.fna"))
    monkeypatch.setattr(estimate, "EST_BITS", 1 << 24)  # the oracle counter is a CPU loop: keep its filter small
    _orig = estimate.estimate_distinct_kmers
    monkeypatch.setattr(estimate, "estimate_distinct_kmers",
                        lambda fp, k, counter, num_bits=1 << 24: _orig(fp, k, counter, num_bits=num_bits))
    db = core.Database("db", str(tmp_path / "db"), str(tmp_path / "tmp"), nproc=1,
                       backend=cpu_backend.OracleBackend(dataset.K, dataset.FILTER_BITS))
    db.set_configs(paths, min_kmer_occurrence=1, kmer_size=dataset.K, filter_expand_by=50.0)
    exact = _exact_distinct(paths, dataset.K, oracle)
    got = db.metadata["filter_size"]
    assert abs(got - 1.5 * exact) / (1.5 * exact) < 0.005
    with open(os.path.join(str(tmp_path / "db"), "metadata.json")) as f:
        assert json.load(f)["filter_size"] == got


@pytest.mark.gpu
@pytest.mark.parametrize("num_bits,s", [(1 << 24, 1), (1 << 20, 4), (1 << 30, 1)])
def test_cuda_counter_matches_oracle_counter(tmp_path, ctx, oracle, num_bits, s):
    from metasbt_b200.backend import CudaBackend
    dataset.write(str(tmp_path / "inputs"))
    paths = sorted(str(p) for p in (tmp_path / "inputs").rglob("*.fna"))
    if num_bits == 1 << 30:
        paths = paths[:3]
    want = cpu_backend.OracleBackend(dataset.K, dataset.FILTER_BITS).count_union_bits(paths, dataset.K, num_bits, s * num_bits)
    got = CudaBackend(dataset.K, dataset.FILTER_BITS).count_union_bits(paths, dataset.K, num_bits, s * num_bits)
    assert got == want and got > 0


@pytest.mark.gpu
def test_set_configs_estimates_on_the_gpu(tmp_path, ctx, oracle):
    """No backend injected: Database.set_configs takes the CUDA counter (2^30-bit counting filter)."""
    dataset.write(str(tmp_path / "inputs"))
    paths = sorted(str(p) for p in (tmp_path / "inputs").rglob("*.fna"))[:4]
    db = core.Database("db", str(tmp_path / "db"), str(tmp_path / "tmp"), nproc=1)
    db.set_configs(paths, min_kmer_occurrence=1, kmer_size=dataset.K, filter_expand_by=50.0)
    exact = _exact_distinct(paths, dataset.K, oracle)
    assert abs(db.metadata["filter_size"] - 1.5 * exact) / (1.5 * exact) < 0.005
