"""Synthetic genomes of the shape BASELINE.json names, generated on the GPU
with plain torch integer ops (input generation only -- never inside a timed region):

    base i of genome g = "ACGT"[splitmix64((seed_g << 32) ^ i) >> 62],  seed_g = 0x5EED0000 + g

and packed into the device layout of msbt_spec.h (32 bases per u64, first base in the top bits).
tests/util.py holds the numpy twin used to cross-check this generator against the oracle.
"""

from __future__ import annotations

import torch

from . import _lib

_C1 = -7046029254386353131  # 0x9E3779B97F4A7C15 as int64
_C2 = -4658895280553007687  # 0xBF58476D1CE4E5B9
_C3 = -7723592293110705685  # 0x94D049BB133111EB


def _lsr(x: torch.Tensor, s: int) -> torch.Tensor:
    """logical shift right on int64"""
    return (x >> s) & ((1 << (64 - s)) - 1)


def splitmix64(x: torch.Tensor) -> torch.Tensor:
    z = x + _C1
    z = (z ^ _lsr(z, 30)) * _C2
    z = (z ^ _lsr(z, 27)) * _C3
    return z ^ _lsr(z, 31)


def genome_codes(seed: int, n: int, device: torch.device) -> torch.Tensor:
    i = torch.arange(n, dtype=torch.int64, device=device)
    return _lsr(splitmix64((seed << 32) ^ i), 62)


def pack_codes_device(codes: torch.Tensor) -> torch.Tensor:
    """int64 codes (0..3) -> packed int64 words (+ PACK_PAD_WORDS zero words)."""
    n = codes.numel()
    n_words = (n + 31) // 32
    padded = torch.zeros(n_words * 32, dtype=torch.int64, device=codes.device)
    padded[:n] = codes
    shifts = torch.arange(62, -2, -2, dtype=torch.int64, device=codes.device)
    words = (padded.view(n_words, 32) << shifts).sum(dim=1)  # disjoint bit fields: sum == or (wraps for the sign bit)
    out = torch.zeros(n_words + _lib.PACK_PAD_WORDS, dtype=torch.int64, device=codes.device)
    out[:n_words] = words
    return out


def packed_genome(seed: int, n: int, device: torch.device) -> torch.Tensor:
    return pack_codes_device(genome_codes(seed, n, device))


def mutated_codes(seed_ancestor: int, seed_mut: int, n: int, per_mille: int, device: torch.device) -> torch.Tensor:
    """Related genomes: the ancestor's bases with SNPs where splitmix64((seed_mut << 32) ^ i) says so."""
    codes = genome_codes(seed_ancestor, n, device)
    r = splitmix64((seed_mut << 32) ^ torch.arange(n, dtype=torch.int64, device=device))
    snp = (_lsr(r, 8) % 1000) < per_mille
    return torch.where(snp, (codes + 1 + _lsr(r, 40) % 3) % 4, codes)


def fasta_text_device(codes: torch.Tensor, name: str, width: int = 80) -> torch.Tensor:
    """uint8 FASTA text (">name\\n" + lines of `width` bases) of a code tensor, assembled on the device -- the bytes a
    FASTA file of this genome would hold (input generation for the end-to-end benchmark, never inside a timed region)."""
    n = codes.numel()
    lut = torch.tensor([65, 67, 71, 84], dtype=torch.uint8, device=codes.device)  # A C G T
    bases = lut[codes]
    full = n // width
    body = torch.full((full, width + 1), 10, dtype=torch.uint8, device=codes.device)  # newline column
    body[:, :width] = bases[: full * width].view(full, width)
    parts = [torch.tensor(list((">" + name + "\n").encode()), dtype=torch.uint8, device=codes.device), body.reshape(-1)]
    if n > full * width:
        parts += [bases[full * width:], torch.tensor([10], dtype=torch.uint8, device=codes.device)]
    return torch.cat(parts)
