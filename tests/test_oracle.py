"""CPU: pin the oracle (a) against MurmurHash3's published vectors and Appleby's reference source, (b) C against its
pure-Python twin, (c) against hand-derived k-mer facts.  PARITY UNPINNED beyond that: the
reference holds no golden vector for this path."""
import os
import random

import numpy as np
import pytest

from tests import util

# MurmurHash3_x64_128 known answers (smhasher reference implementation, seed 0 unless noted)
MURMUR_KAT = [
    (b"", 0, "00000000000000000000000000000000"),
    (b"hello", 0, "cbd8a7b341bd9b025b1e906a48ae1d19"),
    (b"hello, world", 0, "342fac623a5ebc8e4cdcbc079642414d"),
    (b"The quick brown fox jumps over the lazy dog", 0, "e34bbc7bbc071b6c7a433ca9c49a9347"),
]


@pytest.mark.parametrize("data,seed,hexdigest", MURMUR_KAT)
def test_murmur3_known_answers(oracle, data, seed, hexdigest):
    h1, h2 = oracle.murmur3_x64_128(data, seed)
    assert "%016x%016x" % (h1, h2) == hexdigest
    assert oracle.py_murmur3_x64_128(data, seed) == (h1, h2)


def test_murmur3_c_vs_python_all_lengths(oracle):
    rng = random.Random(3)
    for n in range(0, 50):
        data = bytes(rng.randrange(256) for _ in range(n))
        for seed in (0, 1, 0xDEADBEEF):
            assert oracle.murmur3_x64_128(data, seed) == oracle.py_murmur3_x64_128(data, seed)


def _appleby_reference(tmp_path):
    """Austin Appleby's public-domain MurmurHash3.cpp -- the smhasher file HowDeSBT itself vendors -- as shipped in the
    source tree of scikit-learn in this image (third-party code, not part of /root/reference, compiled where it lies)."""
    import ctypes
    import shutil
    import subprocess
    try:
        import sklearn
    except Exception:
        pytest.skip("scikit-learn is not installed")
    src = os.path.join(os.path.dirname(sklearn.__file__), "utils", "src", "MurmurHash3.cpp")
    if not os.path.exists(src) or shutil.which("g++") is None:
        pytest.skip("MurmurHash3.cpp or g++ not available")
    so = str(tmp_path / "libappleby_murmur3.so")
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-I", os.path.dirname(src), "-o", so, src])
    lib = ctypes.CDLL(so)
    fn = getattr(lib, "_Z19MurmurHash3_x64_128PKvijPv", None) or lib.MurmurHash3_x64_128
    fn.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_uint32, ctypes.c_void_p]
    fn.restype = None
    return fn


def test_murmur3_against_appleby_reference_source(oracle, tmp_path):
    """The oracle's MurmurHash3_x64_128 (and through it msbt_spec.h's device copy, which the GPU parity tests tie to the
    oracle) against the original reference implementation, on every key length the k-mer path can produce (k = 1..64 ->
    1..16 bytes) and beyond, with several seeds; plus the exact byte strings the k-mer hash feeds it."""
    import ctypes
    import random
    ref = _appleby_reference(tmp_path)
    rng = random.Random(2024)
    out = (ctypes.c_uint64 * 2)()
    for n in list(range(0, 50)) + [63, 64, 65, 255]:
        for seed in (0, 1, 0x9747B28C, 0xFFFFFFFF):
            data = bytes(rng.getrandbits(8) for _ in range(n))
            ref(data, n, seed, ctypes.addressof(out))
            assert oracle.murmur3_x64_128(data, seed) == (int(out[0]), int(out[1])), (n, seed)
    for k in (1, 4, 5, 21, 31, 32, 33, 51, 64):
        kmer = "".join(rng.choice("ACGT") for _ in range(k))
        canon = oracle.py_canonical(kmer)
        data = canon.to_bytes(16, "little")[:(k + 3) // 4]
        ref(data, len(data), 0, ctypes.addressof(out))
        assert oracle.kmer_hash(kmer, 0) == int(out[0]), (k, kmer)


def test_canonical_rule(oracle):
    # smaller of k-mer and reverse complement under A<C<G<T, first base most significant
    assert oracle.py_canonical("ACGT") == oracle.py_canonical("ACGT")  # palindrome
    assert oracle.py_canonical("TTTT") == 0  # rc = AAAA
    assert oracle.py_canonical("AAAC") == 1
    assert oracle.py_canonical("GTTT") == 1  # rc = AAAC
    for kmer in ("ACGTTGCA", "GATTACA", "T" * 31, "ACGT" * 8, "G" * 33, "ACGTA" * 12):
        rc = kmer[::-1].translate(str.maketrans("ACGT", "TGCA"))
        assert oracle.kmer_hash(kmer) == oracle.kmer_hash(rc) == oracle.py_kmer_hash(kmer)
        assert oracle.kmer_hash(kmer.lower()) == oracle.kmer_hash(kmer)


def test_kmer_hash_is_murmur_of_packed_bytes(oracle):
    # k=4 "AAAC": canonical integer 1 -> one byte 0x01, len (4+3)/4 = 1
    assert oracle.kmer_hash("AAAC") == oracle.murmur3_x64_128(b"\x01", 0)[0]
    # k=31 uses 8 bytes, k=33 nine
    v = oracle.py_canonical("ACGTACGTACGTACGTACGTACGTACGTACG")
    assert oracle.kmer_hash("ACGTACGTACGTACGTACGTACGTACGTACG") == oracle.murmur3_x64_128(v.to_bytes(8, "little"), 0)[0]
    k33 = "ACGTACGTACGTACGTACGTACGTACGTACGTA"
    assert oracle.kmer_hash(k33) == oracle.murmur3_x64_128(oracle.py_canonical(k33).to_bytes(9, "little"), 0)[0]


@pytest.mark.parametrize("k", [1, 5, 9, 21, 31, 32, 33, 51, 64])
@pytest.mark.parametrize("min_abund", [1, 2, 3])
def test_makebf_c_vs_python(oracle, k, min_abund):
    rng = random.Random(100 * k + min_abund)
    fa = util.messy_fasta(rng, 3, 400) + b">dup\n" + b"ACGTTGCAAGGCTTAAGCCGATCGATCGGATATCGCGCTAGCTAGGATCGATCGACTAGCTAGCATCGA" * 3 + b"\n"
    for num_bits, modulus in ((1 << 12, None), (5000, None), (4096, 5003), (5003, 4000)):
        a = oracle.makebf(fa, k, num_bits, min_abund, modulus=modulus)
        b = oracle.py_makebf(fa, k, num_bits, min_abund, modulus=modulus)
        assert np.array_equal(a, b)
        if min_abund == 1 and k <= 32:
            assert np.array_equal(oracle.makebf(fa, k, num_bits, 1, modulus=modulus, rolling=True), a)
    assert np.array_equal(oracle.positions_distinct(fa, k, 1 << 12), oracle.py_positions_distinct(fa, k, 1 << 12))


def test_fasta_semantics(oracle):
    k, m = 5, 1 << 10
    one = oracle.makebf(b">a\nACGTACGTAC\n", k, m)
    # multi-line record == single line; lowercase == uppercase; CRLF tolerated
    assert np.array_equal(oracle.makebf(b">a\nACGTA\nCGTAC\n", k, m), one)
    assert np.array_equal(oracle.makebf(b">a\nacgtacgtac\n", k, m), one)
    assert np.array_equal(oracle.makebf(b">a\r\nACGTA\r\nCGTAC\r\n", k, m), one)
    # N and record boundaries break k-mers: joined-by-N (core.py:1981) == separate records
    two = oracle.makebf(b">a\nACGTACG\n>b\nTTGCAAT\n", k, m)
    assert np.array_equal(oracle.makebf(b">a\nACGTACGNTTGCAAT\n", k, m), two)
    assert not np.array_equal(oracle.makebf(b">a\nACGTACGTTGCAAT\n", k, m), two)
    # shorter than k -> empty filter; empty input -> empty filter
    assert oracle.bf_popcount(oracle.makebf(b">a\nACGT\n", k, m)) == 0
    assert oracle.bf_popcount(oracle.makebf(b"", k, m)) == 0


def test_query_corollary(oracle):
    """App. A6 corollary: with min=1, hits = popc(Q & F) and total = popc(Q)."""
    rng = random.Random(5)
    k, m = 9, 1 << 13
    refs = [util.messy_fasta(rng, 2, 600, "ACGT") for _ in range(4)]
    filters = [oracle.makebf(f, k, m) for f in refs]
    q = refs[1] + util.messy_fasta(rng, 1, 300, "ACGTN")
    pos = oracle.positions_distinct(q, k, m)
    Q = oracle.makebf(q, k, m)
    assert len(pos) == oracle.bf_popcount(Q)
    hits = oracle.query_hits(pos, filters)
    for l, f in enumerate(filters):
        assert hits[l] == oracle.bf_and_popcount(Q, f)
    assert hits[1] >= oracle.bf_popcount(filters[1]) - 0  # the query contains reference 1 entirely
    assert hits[1] == oracle.bf_popcount(filters[1])


def test_filter_algebra(oracle):
    rng = np.random.default_rng(1)
    a, b, c = (rng.integers(0, 1 << 63, size=257, dtype=np.uint64) for _ in range(3))
    assert np.array_equal(oracle.bf_or([a, b, c]), a | b | c)
    pc = lambda x: int(np.unpackbits(x.view(np.uint8)).sum())
    assert oracle.bf_popcount(a) == pc(a)
    assert oracle.bf_and_popcount(a, b) == pc(a & b)
    assert oracle.bf_or_popcount(a, b) == pc(a | b)
    assert oracle.bf_and_popcount(a, b) + oracle.bf_or_popcount(a, b) == pc(a) + pc(b)


def test_ani_formula(oracle):
    # core.py:1816-1829
    assert oracle.ani_distance(0, 10, 31) == 1.0
    assert oracle.ani_distance(10, 10, 31) == 0.0
    d = oracle.ani_distance(5000, 10000, 31)
    assert abs(d - (-(1 / 31) * __import__("math").log(2 * 0.5 / 1.5))) < 1e-15
    assert oracle.ani_distance(1, 10 ** 9, 3) == 1.0  # jaccard rounds to 0.0
    with pytest.raises(ZeroDivisionError):
        oracle.ani_distance(0, 0, 31)


def _kats():
    import json
    import os
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "kat_positions.json")) as f:
        return json.load(f)


@pytest.mark.parametrize("kat", _kats(), ids=lambda c: c["name"])
def test_committed_known_answers(oracle, kat):
    """tests/golden/kat_positions.json (made by tests/golden/make_golden.py): both restatements reproduce it."""
    fa = kat["fasta"].encode()
    for impl in (oracle.makebf, oracle.py_makebf):
        w = impl(fa, kat["k"], kat["num_bits"], kat["min"], modulus=kat["modulus"])
        got = [i * 64 + b for i, x in enumerate(w) for b in range(64) if (int(x) >> b) & 1]
        assert got == kat["set_bits"]
    assert [int(p) for p in oracle.positions_distinct(fa, kat["k"], kat["num_bits"], modulus=kat["modulus"])] == kat["query_positions"]
