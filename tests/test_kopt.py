"""k-mer size search (second half; reference core.py:439-468 shells out to `kitsune kopt`).

kitsune is absent and unpinned, so kopt.py restates the published three-step method (CRE -> ACF -> OFC); the checker is
an independent dictionary/set restatement in oracle/oracle.py (`py_relative_entropy`, `py_kopt`).  CPU: the torch word
counting against the dictionaries, the Bloom-popcount set-size estimators against exact sets, the chosen k, and
Database.set_configs.  GPU: the CUDA backend returns the very integers the oracle backend returns (K1 with modulus
sampling, K3 all pairs, msbt_bf_occurrence), the word counting on the device agrees with the CPU, same k."""
import json
import os

import numpy as np
import pytest

from metasbt_b200 import core, kopt
from tests import cpu_backend, util


def _family(tmp_path, k_species=((0xA11CE010, 6000), (0xA11CE011, 5000)), per_species=3):
    paths, fastas = [], []
    os.makedirs(str(tmp_path), exist_ok=True)
    for sp, (seed, length) in enumerate(k_species):
        ancestor = util.synthetic_codes(seed, length)
        for i in range(per_species):
            codes = util.mutate(ancestor, 0.02 * (1 + i), 100 + sp * 10 + i)
            fa = util.codes_to_fasta(codes, "g%d_%d" % (sp, i), 70)
            p = os.path.join(str(tmp_path), "g%d_%d.fna" % (sp, i))
            with open(p, "wb") as f:
                f.write(fa)
            paths.append(p)
            fastas.append(fa)
    return paths, fastas


MESSY = [
    b">r1 some text\nACGTACGTTTGACCA\nNNNNACGTTGCAAGGCT\r\nacgtaacgttagc\n>r2\nGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGGG\n>r3\n\n>r4\nAC\n",
    b"ACGTTGCATGCATGCCGTAGCTAGCTAGGATCGATCGTAGCTAGCTAGCTAGTCGATGCTAGCTAGTCGATCGATGCTAGCTGATCGTAGCTAGCTAGCTGATCGATCGTAGC",
    b">only header\n",
    b"",
]


def test_fasta_codes_follow_the_record_parser(oracle):
    rng = np.random.default_rng(3)
    seq = "".join("ACGT"[i] for i in rng.integers(0, 4, 900))
    texts = MESSY + [(">a\n" + "\n".join(seq[i:i + 60] for i in range(0, 900, 60)) + "\n").encode()]
    for text in texts:
        codes = kopt.fasta_codes(text)
        recs = oracle.py_fasta_records(text)
        want = []
        for r in recs:
            want.append(4)
            want.extend("ACGT".index(c) if c in "ACGT" else 4 for c in r.upper())
        # the break a header leaves may be missing in front of a headerless first record: compare without leading breaks
        got = list(codes)
        while got and got[0] == 4:
            got.pop(0)
        while want and want[0] == 4:
            want.pop(0)
        assert got == want


def test_relative_entropies_match_the_dictionary_restatement(oracle):
    rng = np.random.default_rng(5)
    seq = "".join("ACGT"[i] for i in rng.integers(0, 4, 2500))
    repeat = seq[:300] * 4 + seq[300:800]
    texts = MESSY[:2] + [
        (">x\n" + seq[:1200] + "\n>y\n" + seq[1200:1700] + "NN" + seq[1700:].lower() + "\n").encode(),
        (">rep\n" + repeat + "\n").encode(),
    ]
    for text in texts:
        got = kopt.relative_entropies(kopt.fasta_codes(text), 32, "cpu")
        for l in list(range(3, 14)) + [20, 31, 32]:
            assert got[l] == pytest.approx(oracle.py_relative_entropy(text, l), abs=1e-9), l
    with pytest.raises(ValueError):
        kopt.relative_entropies(kopt.fasta_codes(texts[0]), 33, "cpu")


def test_limits_and_choice():
    cre = {4: 10.0, 5: 6.0, 6: 2.0, 7: 0.9, 8: 0.1}
    assert kopt.cre_lower_limit(cre) == 7 and kopt.cre_lower_limit(cre, 0.5) == 6
    assert kopt.cre_lower_limit({4: 0.0, 5: 0.0}) == 4
    assert kopt.cre_lower_limit({4: 1.0, 5: 0.9}) == 5            # never below the cut-off: the last k
    acf = {4: 100.0, 5: 500.0, 6: 300.0, 7: 60.0, 8: 40.0}
    assert kopt.acf_upper_limit(acf) == 7 and kopt.acf_upper_limit(acf, 0.5) == 6
    ofc = {4: 1.0, 5: 5.0, 6: 9.0, 7: 9.0, 8: 2.0}
    assert kopt.choose(5, 8, ofc) == 6                             # first of the tie
    assert kopt.choose(7, 8, ofc) == 7
    assert kopt.choose(8, 6, ofc) == 8                             # empty range: the lower limit
    assert kopt.pick_genomes(["b", "a", "c"], 8) == ["a", "b", "c"]
    assert kopt.pick_genomes([str(i).zfill(3) for i in range(100)], 4) == ["000", "025", "050", "075"]


@pytest.mark.parametrize("sample", [1, 3])
def test_set_size_estimators_against_exact_sets(tmp_path, oracle, sample):
    paths, fastas = _family(tmp_path / "fam")
    backend = cpu_backend.OracleBackend(16, 1 << 20)
    m = 1 << 20
    for k in (6, 9, 14):
        sets = [set(oracle.py_word_counts(fa, k)) for fa in fastas]
        pop, inter, ge1, ge2 = backend.kmer_set_profile(paths, k, m, sample * m)
        tol = 0.02 if sample == 1 else 0.12      # sampling noise ~ sqrt(S / n) on sets of a few thousand
        for i, j in ((0, 1), (0, 2), (3, 5)):
            exact = len(sets[i] & sets[j])
            assert abs(kopt.common_features(pop[i], pop[j], inter[i][j], m, sample) - exact) <= tol * exact + 3 * sample
        seen, twice = set(), set()
        for s in sets:
            twice |= seen & s
            seen |= s
        assert abs(kopt.shared_features(ge1, ge2, m, sample) - len(twice)) <= tol * len(twice) + 3 * sample
    # the collision correction matters once the filter is loaded: 6 x ~5,500 14-mers in 2^16 bits (load 0.3)
    small = 1 << 16
    pop, inter, ge1, ge2 = backend.kmer_set_profile(paths, 14, small, small)
    assert abs(kopt.shared_features(ge1, ge2, small, 1) - len(twice)) < 0.05 * len(twice)
    assert ge2 > 1.10 * len(twice)               # what the raw bit count would have claimed


def test_optimal_kmer_size_matches_the_exact_restatement(tmp_path, oracle):
    paths, fastas = _family(tmp_path / "fam")
    exact = oracle.py_kopt(fastas, 14)
    details = {}
    k = kopt.optimal_kmer_size(paths, 14, cpu_backend.OracleBackend(14, 1 << 20), num_bits=1 << 20, details=details)
    assert (k, details["k_lo"], details["k_hi"], details["cre_limits"]) == (exact["k"], exact["k_lo"], exact["k_hi"], exact["cre_limits"])
    assert details["sample"] == 1
    for kk in range(4, 15):
        assert details["acf"][kk] == pytest.approx(exact["acf"][kk], rel=0.02, abs=2)
        assert details["ofc"][kk] == pytest.approx(exact["ofc"][kk], rel=0.02, abs=2)
    # one genome: the CRE limit alone
    assert kopt.optimal_kmer_size(paths[:1], 14, cpu_backend.OracleBackend(14, 1 << 20), num_bits=1 << 20) == exact["cre_limits"][0]
    with pytest.raises(ValueError):
        kopt.optimal_kmer_size(paths, 33, cpu_backend.OracleBackend(14, 1 << 20))
    with pytest.raises(ValueError):
        kopt.optimal_kmer_size([], 14, cpu_backend.OracleBackend(14, 1 << 20))


def test_set_configs_searches_the_kmer_size(tmp_path, oracle, monkeypatch):
    paths, fastas = _family(tmp_path / "fam", k_species=((0xA11CE010, 3000),), per_species=3)
    monkeypatch.setattr(kopt, "SET_BITS", 1 << 18)
    backend = cpu_backend.OracleBackend(12, 1 << 18)
    db = core.Database("db", str(tmp_path / "db"), str(tmp_path / "tmp"), nproc=1, backend=backend)
    with pytest.raises(ValueError):
        db.set_configs(paths, min_kmer_occurrence=1, kmer_size=None, kmer_max=None, filter_size=1 << 18)
    db.set_configs(paths, min_kmer_occurrence=1, kmer_size=None, kmer_max=12, filter_size=1 << 18)
    want = oracle.py_kopt(fastas, 12)["k"]
    assert db.metadata["kmer_size"] == want
    meta = json.load(open(os.path.join(db.root, "metadata.json")))
    assert meta["kmer_size"] == want and meta["filter_size"] == 1 << 18


def test_cli_index_without_kmer_size(tmp_path, oracle, monkeypatch):
    """`metasbt index` without --kmer-size (metasbt.py:441-446: kitsune up to --limit-kmer-size) builds the database at
    the k the search returns."""
    from metasbt_b200 import cli
    from tests import dataset
    refs, _mags, _ = dataset.write(str(tmp_path / "in"))
    paths = sorted(str(p) for p in (tmp_path / "in").rglob("*.fna"))
    monkeypatch.setattr(kopt, "SET_BITS", 1 << 22)
    want = kopt.optimal_kmer_size(paths, 12, cpu_backend.OracleBackend(12, 1 << 22), num_bits=1 << 22)
    assert 8 <= want <= 12
    cli.MetaSBT(["index", "--workdir", str(tmp_path / "w"), "--database", "db", "--references", refs, "--limit-kmer-size", "12",
                 "--filter-size", str(1 << 22), "--min-kmer-occurrences", "1", "--nproc", "1"],
                backend=cpu_backend.OracleBackend(want, 1 << 22))
    meta = json.load(open(str(tmp_path / "w" / "db" / "metadata.json")))
    assert meta["kmer_size"] == want and meta["filter_size"] == 1 << 22
    assert os.path.isfile(str(tmp_path / "w" / "db" / "clusters.tsv"))


# ------------------------------------------------------------------ GPU

@pytest.mark.gpu
def test_bf_occurrence_matches_numpy(ctx):
    import torch
    from metasbt_b200 import ops
    rng = np.random.default_rng(11)
    for num_bits in (64 * 7, 200_003, 1 << 20):
        words = ops.filter_words(num_bits)
        for n in (1, 2, 3, 8, 33):
            host = rng.integers(0, 1 << 63, size=(n, words), dtype=np.int64) & rng.integers(0, 1 << 63, size=(n, words), dtype=np.int64)
            host &= rng.integers(0, 1 << 63, size=(n, words), dtype=np.int64)           # density 1/8
            dev = ctx.alloc_filters(n, num_bits)
            dev.zero_()
            dev[:, :words] = torch.from_numpy(host).to(dev.device)
            got = ctx.bf_occurrence([dev[i] for i in range(n)], num_bits).cpu().tolist()
            once, twice = np.zeros(words, dtype=np.int64), np.zeros(words, dtype=np.int64)
            for row in host:
                twice |= once & row
                once |= row
            want = [int(np.unpackbits(once.view(np.uint8)).sum()), int(np.unpackbits(twice.view(np.uint8)).sum())]
            assert got == want, (num_bits, n)
    assert ctx.bf_occurrence([], 1 << 20).cpu().tolist() == [0, 0]


@pytest.mark.gpu
def test_kmer_set_profile_cuda_equals_the_oracle_backend(tmp_path, oracle):
    from metasbt_b200 import backend as cuda_backend
    paths, _ = _family(tmp_path / "fam")
    with open(paths[0], "ab") as f:                     # a second record, an N run and lowercase in one genome
        f.write(b">extra\nACGTNNNNNNacgtgtgtcagtcgatcgatcgtagctagctagctagctagcatcgatcgatgcatgcatcgatgc\n")
    m = 1 << 20
    cuda = cuda_backend.CudaBackend(31, m)
    cpu = cpu_backend.OracleBackend(31, m)
    before = cuda.ctx.launch_count
    for k, sample in ((4, 1), (11, 1), (21, 3), (32, 1)):
        assert cuda.kmer_set_profile(paths, k, m, sample * m) == cpu.kmer_set_profile(paths, k, m, sample * m), (k, sample)
    assert cuda.ctx.launch_count > before
    cuda.kmer_set_profile_done()


@pytest.mark.gpu
def test_word_counting_on_the_device_and_the_chosen_k(tmp_path, oracle, ctx):
    from metasbt_b200 import backend as cuda_backend
    paths, fastas = _family(tmp_path / "fam")
    codes = kopt.fasta_codes(fastas[0])
    on_cpu = kopt.relative_entropies(codes, 32, "cpu")
    on_gpu = kopt.relative_entropies(codes, 32, ctx.device)
    for l in range(3, 33):
        assert on_gpu[l] == pytest.approx(on_cpu[l], abs=1e-9)
    m = 1 << 20
    d_gpu, d_cpu = {}, {}
    k_gpu = kopt.optimal_kmer_size(paths, 14, cuda_backend.CudaBackend(14, m), num_bits=m, details=d_gpu)
    k_cpu = kopt.optimal_kmer_size(paths, 14, cpu_backend.OracleBackend(14, m), num_bits=m, details=d_cpu)
    assert k_gpu == k_cpu and d_gpu["acf"] == d_cpu["acf"] and d_gpu["ofc"] == d_cpu["ofc"] and d_gpu["cre_limits"] == d_cpu["cre_limits"]


# ------------------------------------------------------------------ properties (hypothesis, CPU)

def _fasta_like():
    from hypothesis import strategies as st
    line = st.text(alphabet="ACGTacgtNn \t", min_size=0, max_size=40)
    header = st.text(alphabet="abc >;_1", min_size=0, max_size=8).map(lambda t: ">" + t.replace("\n", ""))
    return st.lists(st.one_of(line, header), min_size=0, max_size=12).map(lambda ls: "\n".join(ls).encode())


def test_fasta_codes_property(oracle):
    from hypothesis import given, settings

    @settings(max_examples=150, deadline=None)
    @given(_fasta_like())
    def check(text):
        got = [int(c) for c in kopt.fasta_codes(text)]
        want = []
        for r in oracle.py_fasta_records(text):
            want.append(4)
            want.extend("ACGT".index(c) if c in "ACGT" else 4 for c in r.upper())
        while got and got[0] == 4:
            got.pop(0)
        while want and want[0] == 4:
            want.pop(0)
        # runs of breaks are equivalent (a k-mer never spans one): compare with runs collapsed
        def squeeze(v):
            out = []
            for x in v:
                if not (x == 4 and out and out[-1] == 4):
                    out.append(x)
            while out and out[-1] == 4:
                out.pop()
            return out
        assert squeeze(got) == squeeze(want)
    check()


def test_relative_entropy_property(oracle):
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=40, deadline=None)
    @given(st.lists(st.text(alphabet="ACGTN", min_size=0, max_size=60), min_size=1, max_size=4), st.integers(3, 9))
    def check(records, l):
        text = "".join(">r%d\n%s\n" % (i, r) for i, r in enumerate(records)).encode()
        got = kopt.relative_entropies(kopt.fasta_codes(text), l, "cpu")[l]
        assert got == pytest.approx(oracle.py_relative_entropy(text, l), abs=1e-9)
    check()
