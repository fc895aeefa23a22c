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
    return L


def test_library_exports_every_declared_symbol(built):
    header = open(os.path.join(ROOT, "include", "msbt.h")).read()
    declared = set(re.findall(r"\b(msbt_[a-z0-9_A-Z]+)\s*\(", header))
    declared -= {"msbt_genome_desc"}
    assert declared, "no declarations found"
    assert declared == set(built.SYMBOLS)
    assert set(built.exported_symbols()) == declared
    assert built.lib().msbt_abi_version() == 1


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
