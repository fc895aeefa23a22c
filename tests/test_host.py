"""CPU: the C-ABI library loads and exports every symbol of include/msbt.h; the host-side
FASTA packer and the device k-mer code (compiled for the host by tests/emu) agree with the
oracle.  No compute call needs a GPU here."""
import ctypes
import os
import random
import re
import subprocess

import numpy as np
import pytest

from tests import util

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="session")
def built():
    from metasbt_b200 import _lib
    _lib.build()
    return _lib


@pytest.fixture(scope="session")
def emu(built):
    src = os.path.join(ROOT, "tests", "emu", "emu_kmer.cpp")
    so = os.path.join(ROOT, "tests", "emu", "libemu_kmer.so")
    deps = [src, os.path.join(ROOT, "metasbt_b200", "csrc", "msbt_kmer.h"), os.path.join(ROOT, "metasbt_b200", "csrc", "msbt_spec.h")]
    if not os.path.isfile(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["g++", "-O2", "-Wall", "-shared", "-fPIC", "-o", so, src])
    L = ctypes.CDLL(so)
    L.emu_makebf.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32,
                             ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_int]
    L.emu_makebf.restype = ctypes.c_int
    L.emu_makebf_generic.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32,
                                     ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_void_p,
                                     ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32]
    L.emu_makebf_generic.restype = ctypes.c_int
    return L


def test_library_exports_every_declared_symbol(built):
    header = open(os.path.join(ROOT, "include", "msbt.h")).read()
    declared = set(re.findall(r"\b(msbt_[a-z0-9_A-Z]+)\s*\(", header))
    declared -= {"msbt_genome_desc"}
    assert declared, "no declarations found"
    assert declared == set(built.SYMBOLS)
    assert set(built.exported_symbols()) == declared
    assert built.lib().msbt_abi_version() == 2


def test_header_is_plain_c():
    """include/msbt.h is the boundary a cgo / JNI / ctypes binding reads: it must compile as C99 on its own."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.check_call(["gcc", "-fsyntax-only", "-x", "c", "-std=c99", "-Wall", "-Wextra", "-Werror",
                           os.path.join(root, "include", "msbt.h")])


def test_no_torch_types_in_abi():
    header = open(os.path.join(ROOT, "include", "msbt.h")).read()
    assert "torch" not in header.lower().replace("no torch", "") and "at::" not in header and "std::" not in header


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "metasbt_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", text, re.M), f
                assert "msbt_oracle" not in text, f


def _pack_reference(fasta: bytes, k: int):
    """numpy restatement of the packed layout from the oracle's record parser."""
    from oracle import oracle as O
    codes, ends = [], []
    n = 0
    for rec in O.py_fasta_records(fasta):
        run = []
        for ch in rec.upper() + "#":
            if ch in "ACGT":
                run.append("ACGT".index(ch))
            else:
                if len(run) >= max(k, 1):
                    codes.extend(run)
                    n += len(run)
                    ends.append(n)
                run = []
    return np.array(codes, dtype=np.uint8), np.array(ends, dtype=np.uint32)


@pytest.mark.parametrize("k", [1, 4, 31, 33, 64])
def test_fasta_pack_layout(built, k):
    from metasbt_b200 import ops
    rng = random.Random(k)
    for trial in range(30):
        fa = util.messy_fasta(rng, rng.randint(0, 4), rng.choice([5, 80, 500]), rng.choice(["ACGT", "ACGTN", "ACGTacgtNnRY"]))
        pg = ops.pack_fasta(fa, k)
        codes, ends = _pack_reference(fa, k)
        assert pg.n_bases == len(codes)
        assert np.array_equal(pg.run_ends, ends)
        ref = ops.pack_codes(codes)
        assert np.array_equal(pg.words, ref.words)
        assert len(pg.words) == (len(codes) + 31) // 32 + built.PACK_PAD_WORDS


def test_fasta_pack_errors(built):
    L = built.lib()
    nb, nr = ctypes.c_uint64(), ctypes.c_uint64()
    assert L.msbt_fasta_pack(b"ACGT", 4, 65, None, 0, None, 0, ctypes.byref(nb), ctypes.byref(nr)) == built.MSBT_EINVAL
    words = np.zeros(1, dtype=np.uint64)
    ends = np.zeros(1, dtype=np.uint32)
    text = b">x\n" + b"ACGT" * 40 + b"\n"
    assert L.msbt_fasta_pack(text, len(text), 4, words.ctypes.data, 1, ends.ctypes.data, 1, ctypes.byref(nb), ctypes.byref(nr)) == built.MSBT_ENOSPC


@pytest.mark.parametrize("impl", [0, 1], ids=["walker", "block"])
@pytest.mark.parametrize("k", [1, 2, 9, 16, 21, 31, 32, 33, 34, 51, 60, 61, 63, 64])
def test_device_kmer_code_on_host_matches_oracle(emu, oracle, k, impl):
    """The per-thread code the kernels run (walker, run mask, hash, modulus) == oracle makebf."""
    from metasbt_b200 import ops
    rng = random.Random(1000 + k)
    for trial in range(12):
        fa = util.messy_fasta(rng, rng.randint(1, 5), rng.choice([10, 100, 700, 3000]), rng.choice(["ACGT", "ACGTN", "ACGTacgtNnRY"]))
        num_bits = rng.choice([64, 1000, 4096, 12345, (1 << 16), (1 << 16) + 3])
        modulus = rng.choice([num_bits, num_bits, num_bits + 17, max(1, num_bits - 5)])
        pg = ops.pack_fasta(fa, k)
        w = np.zeros((num_bits + 63) // 64, dtype=np.uint64)
        ends = pg.run_ends if len(pg.run_ends) else np.zeros(1, dtype=np.uint32)
        assert emu.emu_makebf(pg.words.ctypes.data, ends.ctypes.data, len(pg.run_ends), pg.n_bases, k, 0, modulus, num_bits, w.ctypes.data, impl) == 0
        assert np.array_equal(w, oracle.makebf(fa, k, num_bits, 1, 0, modulus))


def test_context_requires_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from metasbt_b200 import ops
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.Context(0)


@pytest.mark.parametrize("num_bits", [64, 1000, 4096 * 64, 4096 * 64 + 1, (1 << 20) + 77])
def test_positions_expand_is_the_inverse_of_extraction(built, oracle, num_bits):
    """Host side of the sparse filter transfer (msbt_positions_expand): positions -> the oracle's filter words."""
    from metasbt_b200 import ops
    rng = random.Random(num_bits)
    fa = util.messy_fasta(rng, 3, 4000, "ACGTN")
    want = oracle.makebf(fa, 21, num_bits)
    pos = oracle.positions_distinct(fa, 21, num_bits).astype(np.uint32)
    out = np.full(len(want) + 3, 0xFFFFFFFFFFFFFFFF, dtype=np.uint64)  # stale content must be overwritten, the tail left alone
    ops.positions_expand(pos, num_bits, out)
    assert np.array_equal(out[: len(want)], want) and np.all(out[len(want):] == 0xFFFFFFFFFFFFFFFF)
    ops.positions_expand(np.zeros(0, dtype=np.uint32), num_bits, out)
    assert not out[: len(want)].any()
    with pytest.raises(ValueError):  # beyond the filter
        ops.positions_expand(np.array([num_bits], dtype=np.uint32), num_bits, out)
    if len(pos) > 2 and pos[-1] >= 4096 * 64:
        with pytest.raises(ValueError):  # not ascending across blocks
            ops.positions_expand(pos[::-1].copy(), num_bits, out)


def test_numa_helpers():
    from metasbt_b200 import numa
    assert numa._parse_cpulist("0-3,8,10-11\n") == [0, 1, 2, 3, 8, 10, 11]
    assert numa._parse_cpulist("") == []
    info = numa.bind_to_gpu_node(0)  # no GPU here: nothing is bound, nothing raises
    assert info["bound"] in (False, True)


def test_bench_reference_arm_prints_one_json_line():
    """bench.py --impl reference (the CPU arm the driver runs beside ours) on a tiny configuration."""
    import json
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--genome-bp", "20000", "--filter-bits", str(1 << 20), "--cpu-genomes", "4"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "Mbp/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["e2e"]["h2d_bytes_per_step"] == 0 and d["gpu_launches"] == 0


SPECS = [("quarter", 0, "min"), ("words", 0, "min"), ("quarter", 1, "min"), ("quarter", 0, "forward"), ("words", 1, "forward")]


@pytest.mark.parametrize("key_bytes,digest_half,canonical", SPECS)
def test_hash_spec_switches_oracle_c_vs_python_twin_vs_device_code(emu, oracle, key_bytes, digest_half, canonical):
    """msbt_hash_spec: the unpinned HowDeSBT-side choices as run-time data.  For every switch combination the oracle's C
    restatement, its independent pure-Python twin and the device's generic K1 code (compiled for the host) agree."""
    from metasbt_b200 import ops
    rng = random.Random(hash((key_bytes, digest_half, canonical)) & 0xFFFF)
    try:
        oracle.set_spec(key_bytes, digest_half, canonical)
        for k in (1, 5, 21, 31, 32, 33, 47, 64):
            kmer = "".join(rng.choice("ACGT") for _ in range(k))
            assert oracle.kmer_hash(kmer, 7) == oracle.py_kmer_hash(kmer, 7, key_bytes, digest_half, canonical)
            fa = util.messy_fasta(rng, 3, 400, "ACGTN") + util.codes_to_fasta(util.synthetic_codes(k, 300), "clean")
            num_bits = rng.choice([1000, 1 << 14, (1 << 14) + 3])
            want = oracle.makebf(fa, k, num_bits)
            pg = ops.pack_fasta(fa, k)
            w = np.zeros((num_bits + 63) // 64, dtype=np.uint64)
            ends = pg.run_ends if len(pg.run_ends) else np.zeros(1, dtype=np.uint32)
            assert emu.emu_makebf_generic(pg.words.ctypes.data, ends.ctypes.data, len(pg.run_ends), pg.n_bases, k, 0, num_bits, num_bits,
                                          w.ctypes.data, {"quarter": 0, "words": 1}[key_bytes], digest_half,
                                          {"min": 0, "forward": 1}[canonical]) == 0
            assert np.array_equal(w, want), (k, num_bits)
            same_key = key_bytes == "quarter" or (k + 3) // 4 == (8 if k <= 32 else 16)
            changed = digest_half == 1 or canonical == "forward" or not same_key
            if changed and k >= 5 and want.any():
                oracle.set_spec()
                assert not np.array_equal(oracle.makebf(fa, k, num_bits), want)  # the switch really changes the filter
                oracle.set_spec(key_bytes, digest_half, canonical)
    finally:
        oracle.set_spec()


def test_hash_spec_parsing_and_struct(built):
    from metasbt_b200 import ops
    d = built.HashSpecStruct()
    built.lib().msbt_hash_spec_default(ctypes.byref(d))
    s = ops.HashSpec().to_struct()
    assert (d.struct_size, d.key_bytes, d.digest_half, d.canonical, d.bf_magic, d.bf_version, d.bf_header_size) == \
        (s.struct_size, s.key_bytes, s.digest_half, s.canonical, s.bf_magic, s.bf_version, s.bf_header_size)
    spec = ops.HashSpec.parse("key_bytes=words, digest_half=1,canonical=forward,bf_version=1,bf_magic=0x1234")
    assert spec == ops.HashSpec("words", 1, "forward", 0x1234, 1) and not spec.is_default and ops.HashSpec.parse("").is_default
    for bad in ("key_bytes=sevenths", "digest_half=2", "colour=blue"):
        with pytest.raises(ValueError):
            ops.HashSpec.parse(bad)
    # no GPU here: a context cannot be made, but an invalid spec is rejected before the device is touched
    bad = ops.HashSpec().to_struct()
    bad.key_bytes = 9
    h = ctypes.c_void_p()
    assert built.lib().msbt_ctx_create(0, ctypes.byref(bad), ctypes.byref(h)) == built.MSBT_EINVAL


def test_bf_header_constants_follow_the_spec(tmp_path, monkeypatch):
    from metasbt_b200 import bffile
    words = np.arange(4, dtype=np.uint64)
    old = (bffile.BF_MAGIC, bffile.BF_VERSION)
    try:
        bffile.configure(magic=0x1122334455667788, version=7)
        p = str(tmp_path / "x.bf")
        bffile.write_bf(p, bffile.BloomHeader(kmer_size=21, num_bits=256), words)
        raw = open(p, "rb").read()
        assert int.from_bytes(raw[:8], "little") == 0x1122334455667788 and int.from_bytes(raw[12:16], "little") == 7
        assert np.array_equal(bffile.read_bf(p)[1], words)
        bffile.configure(magic=old[0], version=old[1])
        with pytest.raises(ValueError):
            bffile.read_bf(p)  # another magic: not one of ours
    finally:
        bffile.configure(magic=old[0], version=old[1])


def test_allpairs_route_model():
    """backend.allpairs_route: K3 (AND + popcount per pair) for a handful of filters or dense ones, K4 (transpose + probe with
    the set bits) for many sparse ones."""
    from metasbt_b200 import backend as B
    m = 1 << 28

    def lw(n):
        return ((n + 31) // 32 + 7) // 8 * 8
    sparse = 5_000_000                                       # a 5 Mbp genome: 1.9 % of 2^28 bits
    assert B.allpairs_route(64, m, 64 * sparse, lw(64), 48 << 30) == "bit"
    assert B.allpairs_route(1024, m, 1024 * sparse, lw(1024), 48 << 30) == "probe"
    assert B.allpairs_route(10000, m, 10000 * sparse, lw(10000), 48 << 30) == "probe"
    assert B.allpairs_route(2000, m, 2000 * 60_000_000, lw(2000), 48 << 30) == "bit"     # 22 % dense
    assert B.allpairs_route(1, m, sparse, 8, 48 << 30) == "bit"
    assert B.allpairs_route(5000, 1 << 33, 5000 * sparse, lw(5000), 48 << 30) == "bit"   # positions are 32-bit
    assert B.PROBE_MIN_FILTERS >= 64
