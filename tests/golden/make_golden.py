"""Regenerates tests/golden/config1.json and tests/golden/kat_positions.json with the CPU oracle.

PARITY UNPINNED: the reference holds no golden vector for this path and `howdesbt` is not
available, so these fixtures pin OUR oracle (itself pinned to MurmurHash3's published
vectors and cross-checked C vs pure Python), not a real howdesbt run.  If `howdesbt` is ever on PATH,
tests/pin_against_howdesbt.py re-checks them against the real binary.

    python tests/golden/make_golden.py
"""
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import oracle as O  # noqa: E402
from tests import cpu_backend, dataset, snapshot  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

KAT_CASES = [  # (name, fasta, k, num_bits, modulus, min)
    ("k5_simple", ">a\nACGTACGTAC\n", 5, 1024, None, 1),
    ("k5_lower_multiline", ">a\nacgta\ncgtac\n", 5, 1024, None, 1),
    ("k5_N_breaks", ">a\nACGTACGNTTGCAAT\n", 5, 1024, None, 1),
    ("k5_two_records", ">a\nACGTACG\n>b\nTTGCAAT\n", 5, 1024, None, 1),
    ("k9_palindromes", ">p\nACGTACGTACGTACGTTTTAAAACCCGGGTTTAAACCCGGG\n", 9, 4096, None, 1),
    ("k9_min2", ">p\nACGTTGCAAGGCTTAAGCCACGTTGCAAGGCTTAAGCC\n", 9, 4096, None, 2),
    ("k21_modulus", ">m\nGATTACAGATTACAGATTACACCCTTTGGGAAACATCATCATGGGTTTAAACCCGGGTTT\n", 21, 5000, 4999, 1),
    ("k31", ">m\nGATTACAGATTACAGATTACACCCTTTGGGAAACATCATCATGGGTTTAAACCCGGGTTTACGATCGACTAGCTAGCATCG\n", 31, 1 << 16, None, 1),
    ("k32", ">m\nGATTACAGATTACAGATTACACCCTTTGGGAAACATCATCATGGGTTTAAACCCGGGTTTACGATCGACTAGCTAGCATCG\n", 32, 1 << 16, None, 1),
    ("k33", ">m\nGATTACAGATTACAGATTACACCCTTTGGGAAACATCATCATGGGTTTAAACCCGGGTTTACGATCGACTAGCTAGCATCG\n", 33, 1 << 16, None, 1),
    ("k51", ">m\nGATTACAGATTACAGATTACACCCTTTGGGAAACATCATCATGGGTTTAAACCCGGGTTTACGATCGACTAGCTAGCATCG\n", 51, 1 << 16, None, 1),
    ("k64", ">m\nGATTACAGATTACAGATTACACCCTTTGGGAAACATCATCATGGGTTTAAACCCGGGTTTACGATCGACTAGCTAGCATCG\n", 64, 1 << 16, None, 1),
    ("shorter_than_k", ">s\nACGT\n", 5, 1024, None, 1),
    ("empty", "", 5, 1024, None, 1),
]


def main():
    O.build()
    kats = []
    for name, fasta, k, bits, modulus, mn in KAT_CASES:
        w = O.makebf(fasta.encode(), k, bits, mn, modulus=modulus)
        assert (w == O.py_makebf(fasta.encode(), k, bits, mn, modulus=modulus)).all()
        set_bits = [int(i * 64 + b) for i, x in enumerate(w) for b in range(64) if (int(x) >> b) & 1]
        pos = [int(p) for p in O.positions_distinct(fasta.encode(), k, bits, modulus=modulus)]
        kats.append({"name": name, "fasta": fasta, "k": k, "num_bits": bits, "modulus": modulus, "min": mn,
                     "set_bits": set_bits, "query_positions": pos})
    with open(os.path.join(HERE, "kat_positions.json"), "w") as f:
        json.dump(kats, f, indent=0)
    with tempfile.TemporaryDirectory() as tmp:
        snap = snapshot.run_config1(tmp, backend=cpu_backend.OracleBackend(dataset.K, dataset.FILTER_BITS))
    with open(os.path.join(HERE, "config1.json"), "w") as f:
        json.dump(snap, f, indent=0, sort_keys=True)
    print("wrote", len(kats), "KATs and the config-1 snapshot:", len(snap["filters"]), "filters,", len(snap["profiles"]), "profiles")


if __name__ == "__main__":
    main()
