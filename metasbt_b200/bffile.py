"""Reader / writer of HowDeSBT `.bf` Bloom filter files (simple, uncompressed, one vector).

Layout follows SURVEY.md Appendix A2 (constants live in csrc/msbt_spec.h and are mirrored
here; tests/test_host.py checks that the two agree):

    0x00 u64 magic            0x08 u32 headerSize (0x70)   0x0C u32 version
    0x10 u32 bfKind (simple)  0x14 u32 padding
    0x18 u32 kmerSize         0x1C u32 numHashes
    0x20 u64 hashSeed1        0x28 u64 hashSeed2
    0x30 u64 hashModulus      0x38 u64 numBits
    0x40 u32 numVectors (1)   0x44 u32 setSizeKnown        0x48 u64 setSize
    0x50 vector info: u32 compressor (uncompressed), u32 name offset, u64 data offset,
                      u64 numBytes, u64 filterInfo
    0x70 payload: sdsl::bit_vector = u64 bit length, then ceil(numBits/64) little-endian u64 words

The payload (bit i = (word[i >> 6] >> (i & 63)) & 1) is the part the device kernels read and
write; it is byte-identical to the device layout, so reading is a zero-copy numpy view.
The header is "unpinned" (no real howdesbt file was available): the reader only
relies on magic, headerSize, the (data offset, numBytes) indirection and the fields it returns.
"""

from __future__ import annotations

import os
import struct
from dataclasses import dataclass
from typing import Optional

import numpy as np

BF_MAGIC = 0xD532006662544253
BF_HEADER_SIZE = 0x70
BF_VERSION = 2
BF_KIND_SIMPLE = 1
BF_COMP_UNCOMPRESSED = 1

_HDR = struct.Struct("<QIIIIIIQQQQIIQIIQQQ")  # 0x70 bytes
assert _HDR.size == BF_HEADER_SIZE


def configure(magic: Optional[int] = None, version: Optional[int] = None) -> None:
    """Re-pin the unpinned header constants (ops.HashSpec.bf_magic / bf_version, e.g. from MSBT_HASH_SPEC) without
    touching the code: files are written with, and expected to carry, these values from now on."""
    global BF_MAGIC, BF_VERSION
    if magic is not None:
        BF_MAGIC = int(magic)
    if version is not None:
        BF_VERSION = int(version)


def _configure_from_env() -> None:
    text = os.environ.get("MSBT_HASH_SPEC", "")
    for item in text.split(","):
        key, _, value = item.partition("=")
        if key.strip() == "bf_magic":
            configure(magic=int(value, 0))
        elif key.strip() == "bf_version":
            configure(version=int(value, 0))


_configure_from_env()


@dataclass
class BloomHeader:
    kmer_size: int
    num_bits: int
    num_hashes: int = 1
    seed1: int = 0
    seed2: int = 0
    modulus: Optional[int] = None
    set_size_known: int = 0
    set_size: int = 0

    def __post_init__(self):
        if self.modulus is None:
            self.modulus = self.num_bits


def payload_words(num_bits: int) -> int:
    return (num_bits + 63) // 64


def file_size(num_bits: int) -> int:
    return BF_HEADER_SIZE + 8 + 8 * payload_words(num_bits)


def write_bf(path: str, header: BloomHeader, words: np.ndarray) -> None:
    """Write one filter.  `words` is a uint64 (or int64) array holding at least payload_words() entries."""
    nw = payload_words(header.num_bits)
    w = np.ascontiguousarray(words).view(np.uint64)[:nw]
    if len(w) != nw:
        raise ValueError(f"filter has {len(w)} words, {nw} needed for {header.num_bits} bits")
    num_bytes = 8 + 8 * nw
    hdr = _HDR.pack(BF_MAGIC, BF_HEADER_SIZE, BF_VERSION, BF_KIND_SIMPLE, 0,
                    header.kmer_size, header.num_hashes, header.seed1, header.seed2, header.modulus, header.num_bits,
                    1, header.set_size_known, header.set_size,
                    BF_COMP_UNCOMPRESSED, 0, BF_HEADER_SIZE, num_bytes, 0)
    tmp = f"{path}.tmp{os.getpid()}"
    with open(tmp, "wb") as f:
        f.write(hdr)
        f.write(struct.pack("<Q", header.num_bits))
        f.write(w.tobytes())
    os.replace(tmp, path)


def read_header(path: str) -> BloomHeader:
    with open(path, "rb") as f:
        raw = f.read(BF_HEADER_SIZE)
    return _parse_header(raw, path)[0]


def _parse_header(raw: bytes, path: str):
    if len(raw) < BF_HEADER_SIZE:
        raise ValueError(f"{path}: too short for a .bf header")
    (magic, hsize, _version, kind, _pad, k, nh, s1, s2, mod, nbits, nvec, ssk, ss,
     comp, _name, offset, nbytes, _finfo) = _HDR.unpack(raw[:BF_HEADER_SIZE])
    if magic != BF_MAGIC:
        raise ValueError(f"{path}: not a HowDeSBT bloom filter file (bad magic {magic:#x})")
    if nvec != 1 or kind != BF_KIND_SIMPLE or comp != BF_COMP_UNCOMPRESSED:
        raise ValueError(f"{path}: only simple uncompressed single-vector filters are supported "
                         f"(kind={kind}, vectors={nvec}, compressor={comp})")
    return BloomHeader(k, nbits, nh, s1, s2, mod, ssk, ss), offset if offset else hsize, nbytes


def read_bf(path: str):
    """-> (BloomHeader, uint64 words).  The words array is a fresh copy the caller may pin or upload."""
    with open(path, "rb") as f:
        raw = f.read()
    header, offset, nbytes = _parse_header(raw, path)
    nw = payload_words(header.num_bits)
    (bit_len,) = struct.unpack_from("<Q", raw, offset)
    if bit_len != header.num_bits or nbytes < 8 + 8 * nw or len(raw) < offset + 8 + 8 * nw:
        raise ValueError(f"{path}: payload does not match the header ({bit_len} bits, {nbytes} bytes)")
    words = np.frombuffer(raw, dtype="<u8", count=nw, offset=offset + 8).astype(np.uint64)
    return header, words
