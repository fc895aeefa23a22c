"""Shared helpers for the test-suite: seeded synthetic FASTA generators."""
import random

import numpy as np


def splitmix64(x: np.ndarray) -> np.ndarray:
    x = (x + np.uint64(0x9E3779B97F4A7C15)).astype(np.uint64)
    z = x
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def synthetic_codes(seed: int, n: int) -> np.ndarray:
    """Base codes of a synthetic genome: base i = splitmix64((seed << 32) ^ i) >> 62."""
    with np.errstate(over="ignore"):
        i = np.arange(n, dtype=np.uint64)
        return (splitmix64((np.uint64(seed) << np.uint64(32)) ^ i) >> np.uint64(62)).astype(np.uint8)


def codes_to_fasta(codes: np.ndarray, name: str = "seq", width: int = 80) -> bytes:
    s = np.frombuffer(b"ACGT", dtype=np.uint8)[codes].tobytes()
    lines = [b">" + name.encode()]
    if width <= 0:
        lines.append(s)
    else:
        lines.extend(s[i:i + width] for i in range(0, len(s), width))
    return b"\n".join(lines) + b"\n"


def messy_fasta(rng: random.Random, nrec: int, maxlen: int, alphabet: str = "ACGTacgtNnRY", width: int = 60) -> bytes:
    out = []
    for r in range(nrec):
        n = rng.randint(0, maxlen)
        s = "".join(rng.choice(alphabet if rng.random() < 0.7 else "ACGT") for _ in range(n))
        out.append(">rec%d some description\n" % r)
        for i in range(0, len(s), width):
            out.append(s[i:i + width] + "\n")
    return "".join(out).encode()


def mutate(codes: np.ndarray, rate: float, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    out = codes.copy()
    hit = rng.random(len(codes)) < rate
    out[hit] = (out[hit] + rng.integers(1, 4, size=int(hit.sum()), dtype=np.uint8)) % 4
    return out
