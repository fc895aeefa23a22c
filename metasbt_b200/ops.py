"""Host-side operators over libmsbt.so: one method per `howdesbt` subcommand MetaSBT spawns.

torch is used for device memory and streams only; every computation happens inside the
sm_100a kernels of libmsbt.so, reached through the C-ABI of include/msbt.h.  There is no
CPU path: constructing a Context without a CUDA device raises.

Reference call sites replaced (all /root/reference/metasbt/core.py):
    Context.build_filters   howdesbt makebf       :3920-3931
    Context.bf_or           howdesbt bfoperate    :3851-3857
    Context.bf_popcount     howdesbt dumpbf       :3588-3593
    Context.bf_dist         howdesbt bfdistance   :1747-1753 + :1768-1774
    Context.bf_allpairs     the dist loop of Entry.get_boundaries :4036-4081
    Context.query           howdesbt query        :2039-2047
"""

from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib

WORD_BITS = 64


@dataclass(frozen=True)
class HashSpec:
    """msbt_hash_spec (include/msbt.h): the HowDeSBT-side conventions that are not pinned against a real binary, as data.
    The default is SURVEY.md App. A as written.  A different spec is a configuration change, not a rebuild:
    pass it to Context(...) or set MSBT_HASH_SPEC, e.g. "key_bytes=words,digest_half=1,canonical=forward,bf_version=1"."""
    key_bytes: str = "quarter"     # "quarter": (k+3)/4 bytes of the packed k-mer; "words": whole 64-bit words
    digest_half: int = 0           # which 64-bit half of MurmurHash3_x64_128 is the hash value
    canonical: str = "min"         # "min": min(k-mer, reverse complement), first base most significant; "forward": as read
    bf_magic: int = 0xD532006662544253
    bf_version: int = 2
    bf_header_size: int = 0x70

    @classmethod
    def parse(cls, text: str) -> "HashSpec":
        kw = {}
        for item in filter(None, (t.strip() for t in text.split(","))):
            key, _, value = item.partition("=")
            key, value = key.strip(), value.strip()
            if key in ("key_bytes", "canonical"):
                kw[key] = value
            elif key in ("digest_half", "bf_magic", "bf_version", "bf_header_size"):
                kw[key] = int(value, 0)
            else:
                raise ValueError(f"MSBT_HASH_SPEC: unknown field {key!r}")
        spec = cls(**kw)
        spec.to_struct()
        return spec

    def to_struct(self) -> "_lib.HashSpecStruct":
        if self.key_bytes not in ("quarter", "words") or self.canonical not in ("min", "forward") or self.digest_half not in (0, 1):
            raise ValueError(f"invalid hash spec: {self}")
        st = _lib.HashSpecStruct()
        st.struct_size = ctypes.sizeof(_lib.HashSpecStruct)
        st.key_bytes = 0 if self.key_bytes == "quarter" else 1
        st.digest_half = self.digest_half
        st.canonical = 0 if self.canonical == "min" else 1
        st.bf_magic, st.bf_version, st.bf_header_size = self.bf_magic, self.bf_version, self.bf_header_size
        return st

    @property
    def is_default(self) -> bool:
        return self == HashSpec()


def env_hash_spec() -> Optional[HashSpec]:
    import os
    text = os.environ.get("MSBT_HASH_SPEC", "").strip()
    return HashSpec.parse(text) if text else None


def filter_words(num_bits: int) -> int:
    return (num_bits + 63) // 64


def filter_stride_words(num_bits: int) -> int:
    """Row stride of a filter matrix: words rounded up to 16 bytes so every row is uint4-aligned."""
    w = filter_words(num_bits)
    return (w + 1) // 2 * 2


@dataclass
class PackedGenome:
    """2-bit packed valid bases of one FASTA plus its run table (include/msbt.h: msbt_fasta_pack)."""
    words: np.ndarray      # uint64, ceil(n_bases/32) + PACK_PAD_WORDS
    run_ends: np.ndarray   # uint32 cumulative run ends
    n_bases: int


def pack_fasta(text: bytes, min_run: int) -> PackedGenome:
    """FASTA text -> PackedGenome on the host (C++ in libmsbt.so; no GPU involved).
    Runs shorter than `min_run` (pass k) cannot hold a k-mer and are dropped."""
    L = _lib.lib()
    nb, nr = ctypes.c_uint64(0), ctypes.c_uint64(0)
    rc = L.msbt_fasta_pack(text, len(text), min_run, None, 0, None, 0, ctypes.byref(nb), ctypes.byref(nr))
    if rc != 0:
        raise ValueError(f"msbt_fasta_pack failed with code {rc}")
    n_words = (nb.value + 31) // 32 + _lib.PACK_PAD_WORDS
    words = np.zeros(n_words, dtype=np.uint64)
    ends = np.zeros(max(nr.value, 1), dtype=np.uint32)
    rc = L.msbt_fasta_pack(text, len(text), min_run, words.ctypes.data, n_words, ends.ctypes.data, len(ends),
                           ctypes.byref(nb), ctypes.byref(nr))
    if rc != 0:
        raise ValueError(f"msbt_fasta_pack failed with code {rc}")
    return PackedGenome(words, ends[: nr.value], int(nb.value))


def positions_expand(positions: np.ndarray, num_bits: int, out: np.ndarray) -> np.ndarray:
    """Host side of a sparse filter transfer (msbt_positions_expand): ascending uint32 positions -> dense filter words
    written into `out` (uint64, >= filter_words(num_bits) entries).  Releases the GIL: callers fan it out over threads."""
    pos = np.ascontiguousarray(positions).view(np.uint32)
    w = filter_words(num_bits)
    if out.dtype.itemsize != 8 or out.size < w or not out.flags["C_CONTIGUOUS"]:
        raise ValueError("positions_expand: `out` must be a contiguous 64-bit array of at least filter_words(num_bits) entries")
    rc = _lib.lib().msbt_positions_expand(pos.ctypes.data if pos.size else None, pos.size, num_bits, out.ctypes.data)
    if rc != 0:
        raise ValueError(f"msbt_positions_expand failed ({rc}): positions must be ascending and below num_bits")
    return out


def pack_codes(codes: np.ndarray) -> PackedGenome:
    """Pack an array of base codes (0..3, one run) -- used for synthetic genomes."""
    codes = np.ascontiguousarray(codes, dtype=np.uint8)
    n = len(codes)
    n_words = (n + 31) // 32
    padded = np.zeros(n_words * 32, dtype=np.uint8)
    padded[:n] = codes
    q = padded.reshape(-1, 4)
    by = (q[:, 0] << 6) | (q[:, 1] << 4) | (q[:, 2] << 2) | q[:, 3]
    words = np.zeros(n_words + _lib.PACK_PAD_WORDS, dtype=np.uint64)
    words[:n_words] = by.astype(np.uint8).view(">u8").astype(np.uint64)
    ends = np.array([n] if n else [], dtype=np.uint32)
    return PackedGenome(words, ends, n)


class GenomeBatch:
    """A batch of packed genomes resident in HBM: one words tensor, one run table, one desc table."""

    def __init__(self, genomes: Sequence[PackedGenome], device: torch.device, filter_indices: Optional[Sequence[int]] = None,
                 pinned: bool = True):
        self.n = len(genomes)
        descs = (_lib.GenomeDesc * max(self.n, 1))()
        woff = roff = 0
        for i, g in enumerate(genomes):
            descs[i].packed_word_off = woff
            descs[i].n_bases = g.n_bases
            descs[i].run_off = roff
            descs[i].n_runs = len(g.run_ends)
            descs[i].filter_index = i if filter_indices is None else filter_indices[i]
            woff += len(g.words)
            roff += len(g.run_ends)
        self.descs = descs
        self.total_bases = sum(g.n_bases for g in genomes)
        h_words = torch.empty(max(woff, 1), dtype=torch.int64, pin_memory=pinned)
        h_runs = torch.empty(max(roff, 1), dtype=torch.int32, pin_memory=pinned)
        wv, rv = h_words.numpy().view(np.uint64), h_runs.numpy().view(np.uint32)
        woff = roff = 0
        for g in genomes:
            wv[woff: woff + len(g.words)] = g.words
            rv[roff: roff + len(g.run_ends)] = g.run_ends
            woff += len(g.words)
            roff += len(g.run_ends)
        self.h_words, self.h_runs = h_words, h_runs
        self.h2d_bytes = h_words.numel() * 8 + h_runs.numel() * 4
        self.device = device
        self.d_words = None
        self.d_runs = None

    @classmethod
    def from_device_words(cls, words: Sequence[torch.Tensor], n_bases: Sequence[int], device: torch.device,
                          keep_host_copy: bool = False) -> "GenomeBatch":
        """Batch of single-run genomes whose packed words already live on the device (synthetic inputs)."""
        self = cls.__new__(cls)
        self.n = len(words)
        descs = (_lib.GenomeDesc * max(self.n, 1))()
        woff = 0
        for i, (w, nb) in enumerate(zip(words, n_bases)):
            descs[i].packed_word_off = woff
            descs[i].n_bases = nb
            descs[i].run_off = i
            descs[i].n_runs = 1 if nb else 0
            descs[i].filter_index = i
            woff += w.numel()
        self.descs = descs
        self.total_bases = int(sum(n_bases))
        self.device = device
        self.d_words = torch.cat(list(words)) if self.n else torch.zeros(1, dtype=torch.int64, device=device)
        self.d_runs = torch.tensor(list(n_bases) or [0], dtype=torch.int64).to(torch.int32).to(device)
        self.h_words = self.h_runs = None
        if keep_host_copy:
            self.h_words = torch.empty(self.d_words.shape, dtype=torch.int64, pin_memory=True).copy_(self.d_words)
            self.h_runs = torch.empty(self.d_runs.shape, dtype=torch.int32, pin_memory=True).copy_(self.d_runs)
        self.h2d_bytes = self.d_words.numel() * 8 + self.d_runs.numel() * 4
        return self

    @classmethod
    def from_device_pack(cls, d_words: torch.Tensor, d_runs: torch.Tensor, descs, n: int, device: torch.device,
                         h2d_bytes: int = 0) -> "GenomeBatch":
        """Batch whose packed words / run table were produced on the device (Context.pack_fasta_device)."""
        self = cls.__new__(cls)
        self.n = n
        self.descs = descs
        self.total_bases = int(sum(descs[i].n_bases for i in range(n)))
        self.device = device
        self.d_words, self.d_runs = d_words, d_runs
        self.h_words = self.h_runs = None
        self.h2d_bytes = h2d_bytes
        return self

    def to_device(self, non_blocking: bool = True) -> "GenomeBatch":
        self.d_words = self.h_words.to(self.device, non_blocking=non_blocking)
        self.d_runs = self.h_runs.to(self.device, non_blocking=non_blocking)
        return self


def _ptr_array(tensors: Sequence[torch.Tensor]):
    arr = (ctypes.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = t.data_ptr()
    return arr


class Context:
    """One msbt_ctx (one GPU).  All methods are asynchronous on torch's current stream."""

    def __init__(self, device: int | torch.device | None = None, spec: Optional[HashSpec] = None):
        if not torch.cuda.is_available():
            raise RuntimeError("metasbt_b200 needs a CUDA device: the hot path has no CPU fallback")
        if device is None:
            device = torch.cuda.current_device()
        self.device = torch.device("cuda", device) if isinstance(device, int) else torch.device(device)
        self._L = _lib.lib()
        h = ctypes.c_void_p()
        self.spec = spec if spec is not None else (env_hash_spec() or HashSpec())
        st = self.spec.to_struct()
        rc = self._L.msbt_ctx_create(self.device.index or 0, None if self.spec.is_default else ctypes.byref(st), ctypes.byref(h))
        if rc != 0:
            raise RuntimeError(f"msbt_ctx_create(device={self.device.index}) failed with code {rc}")
        self._h = h

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._L.msbt_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ helpers
    def _stream(self) -> ctypes.c_void_p:
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _check(self, rc: int, what: str) -> None:
        if rc != 0:
            msg = self._L.msbt_last_error(self._h).decode(errors="replace")
            exc = ValueError if rc == _lib.MSBT_EINVAL else (MemoryError if rc == _lib.MSBT_ENOMEM else RuntimeError)
            raise exc(f"{what} failed ({rc}): {msg}")

    @property
    def launch_count(self) -> int:
        return int(self._L.msbt_launch_count(self._h))

    def sync(self) -> None:
        self._check(self._L.msbt_sync(self._h, self._stream()), "msbt_sync")

    def alloc_filters(self, n: int, num_bits: int) -> torch.Tensor:
        """Uninitialised [n, stride] int64 matrix; row i is filter i (first filter_words(num_bits) words)."""
        return torch.empty((n, filter_stride_words(num_bits)), dtype=torch.int64, device=self.device)

    # ------------------------------------------------------------------ FASTA ingest on the device (N1)
    def pack_fasta_device(self, texts: Sequence[bytes], min_run: int) -> GenomeBatch:
        """FASTA texts (one per genome) -> GenomeBatch, parsed and 2-bit packed ON THE DEVICE (msbt_fasta_pack_device):
        the raw text is copied host -> device once and never touched by the CPU.  Same result as pack_fasta()."""
        n = len(texts)
        sizes = [len(t) for t in texts]
        offs = (ctypes.c_uint64 * (n + 1))()
        for i, sz in enumerate(sizes):
            offs[i + 1] = offs[i] + sz
        total = int(offs[n])
        h_text = torch.empty(max(total, 1), dtype=torch.uint8, pin_memory=True)
        hv = h_text.numpy()
        for i, t in enumerate(texts):
            if sizes[i]:
                hv[offs[i]: offs[i + 1]] = np.frombuffer(t, dtype=np.uint8)
        d_text = h_text.to(self.device, non_blocking=True)
        return self.pack_device_text(d_text, sizes, min_run)

    def pack_device_text(self, d_text: torch.Tensor, sizes: Sequence[int], min_run: int) -> GenomeBatch:
        """The same for FASTA text that already sits on the device: `d_text` (uint8) holds the files back to back,
        `sizes` their lengths in bytes."""
        n = len(sizes)
        offs = (ctypes.c_uint64 * (n + 1))()
        for i, sz in enumerate(sizes):
            offs[i + 1] = offs[i] + int(sz)
        total = int(offs[n])
        if d_text.numel() < total:
            raise ValueError("pack_device_text: the text tensor is shorter than the summed file sizes")
        need = max(int(min_run), 1)
        n_words = sum((sz + 31) // 32 + _lib.PACK_PAD_WORDS for sz in sizes)
        n_runs = sum(sz // need + 1 for sz in sizes)
        d_words = torch.empty(max(n_words, 1), dtype=torch.int64, device=self.device)
        d_runs = torch.empty(max(n_runs, 1), dtype=torch.int32, device=self.device)
        descs = (_lib.GenomeDesc * max(n, 1))()
        rc = self._L.msbt_fasta_pack_device(self._h, d_text.data_ptr(), offs, n, int(min_run), d_words.data_ptr(), d_words.numel(),
                                            d_runs.data_ptr(), d_runs.numel(), descs, self._stream())
        self._check(rc, "msbt_fasta_pack_device")
        return GenomeBatch.from_device_pack(d_words, d_runs, descs, n, self.device, h2d_bytes=total)

    # ------------------------------------------------------------------ K1
    def build_filters(self, batch: GenomeBatch, k: int, num_bits: int, min_abund: int = 1, seed: int = 0,
                      modulus: Optional[int] = None, out: Optional[torch.Tensor] = None, variant: int = 0) -> torch.Tensor:
        """howdesbt makebf --k --min --bits --hashes=1 --seed=seed,0 for every genome of the batch."""
        if batch.d_words is None:
            batch.to_device()
        n_out = (max(d.filter_index for d in batch.descs[: batch.n]) + 1) if batch.n else 0
        if out is None:
            out = self.alloc_filters(n_out, num_bits)
        if out.dtype != torch.int64 or out.dim() != 2 or out.stride(1) != 1 or out.shape[0] < n_out:
            raise ValueError("out must be a [n, stride] int64 matrix with contiguous rows")
        modulus = num_bits if modulus is None else modulus
        rc = self._L.msbt_kmer_bloom_build(
            self._h, batch.d_words.data_ptr(), batch.d_runs.data_ptr(), batch.descs, batch.n, k, seed, modulus,
            num_bits, min_abund, out.data_ptr(), out.stride(0), variant, self._stream())
        self._check(rc, "msbt_kmer_bloom_build")
        return out

    # ------------------------------------------------------------------ K2
    def bf_or(self, filters: Sequence[torch.Tensor], num_bits: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """howdesbt bfoperate --or: OR of the given filter rows."""
        words = filter_words(num_bits)
        if out is None:
            out = torch.empty(filter_stride_words(num_bits), dtype=torch.int64, device=self.device)
        ptrs = _ptr_array(filters)
        rc = self._L.msbt_bf_or_reduce(self._h, ptrs, len(filters), words, out.data_ptr(), self._stream())
        self._check(rc, "msbt_bf_or_reduce")
        return out

    # ------------------------------------------------------------------ K3'
    def bf_popcount(self, filters: Sequence[torch.Tensor], num_bits: int) -> torch.Tensor:
        """howdesbt dumpbf --show:density numerators: popcount of every filter -> int64[n]."""
        out = torch.empty(len(filters), dtype=torch.int64, device=self.device)
        if len(filters):
            rc = self._L.msbt_bf_popcount(self._h, _ptr_array(filters), len(filters), filter_words(num_bits),
                                          out.data_ptr(), self._stream())
            self._check(rc, "msbt_bf_popcount")
        return out

    def bf_occurrence(self, filters: Sequence[torch.Tensor], num_bits: int) -> torch.Tensor:
        """int64[2]: bits set in at least one / in at least two of the filters (one streaming pass; kopt.py)."""
        out = torch.empty(2, dtype=torch.int64, device=self.device)
        rc = self._L.msbt_bf_occurrence(self._h, _ptr_array(filters) if len(filters) else None, len(filters),
                                        filter_words(num_bits), out.data_ptr(), self._stream())
        self._check(rc, "msbt_bf_occurrence")
        return out

    # ------------------------------------------------------------------ K3
    def bf_dist(self, focus: torch.Tensor, targets: Sequence[torch.Tensor], num_bits: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """howdesbt bfdistance --show:intersect and --show:union in one pass -> (I[n], U[n]) int64."""
        n = len(targets)
        inter = torch.empty(n, dtype=torch.int64, device=self.device)
        union = torch.empty(n, dtype=torch.int64, device=self.device)
        if n:
            rc = self._L.msbt_bf_and_popcount_1vN(self._h, focus.data_ptr(), _ptr_array(targets), n,
                                                  filter_words(num_bits), inter.data_ptr(), union.data_ptr(), self._stream())
            self._check(rc, "msbt_bf_and_popcount_1vN")
        return inter, union

    def bf_dist_pairs(self, a: Sequence[torch.Tensor], b: Sequence[torch.Tensor], num_bits: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """(popc(a_p & b_p), popc(a_p | b_p)) for an arbitrary list of pairs in ONE launch -> (I[n], U[n]) int64."""
        n = len(a)
        if len(b) != n:
            raise ValueError("bf_dist_pairs: the two lists must have the same length")
        inter = torch.empty(n, dtype=torch.int64, device=self.device)
        union = torch.empty(n, dtype=torch.int64, device=self.device)
        if n:
            rc = self._L.msbt_bf_and_popcount_pairs(self._h, _ptr_array(a), _ptr_array(b), n, filter_words(num_bits),
                                                    inter.data_ptr(), union.data_ptr(), self._stream())
            self._check(rc, "msbt_bf_and_popcount_pairs")
        return inter, union

    def bf_allpairs(self, filters: Sequence[torch.Tensor], num_bits: int) -> torch.Tensor:
        """Dense n x n intersection matrix (diagonal = popcounts)."""
        n = len(filters)
        out = torch.empty((n, n), dtype=torch.int64, device=self.device)
        if n:
            rc = self._L.msbt_bf_and_popcount_allpairs(self._h, _ptr_array(filters), n, filter_words(num_bits),
                                                       out.data_ptr(), self._stream())
            self._check(rc, "msbt_bf_and_popcount_allpairs")
        return out

    def bf_allpairs_probe(self, filters: Sequence[torch.Tensor], num_bits: int, slice_bytes: int = 48 << 30,
                          batch: int = 256) -> torch.Tensor:
        """The matrix bf_allpairs() returns (intersections, diagonal = popcounts) computed through K4 instead of K3: all
        filters are transposed into a bit-sliced matrix and every filter's set-bit positions probe it, so the work grows
        with the number of SET bits (n * popcount rows of n bits) instead of n^2 / 2 * num_bits -- for sparse filters
        (a 5 Mbp genome in 2^28 bits: 1.9 %) and many of them that is several times less.  The matrix is built for
        `slice_bytes` worth of rows at a time; queries go in batches of `batch` filters.  Same integers as bf_allpairs."""
        n = len(filters)
        out = torch.zeros((n, n), dtype=torch.int64, device=self.device)
        if n == 0:
            return out
        if num_bits > (1 << 32):
            raise ValueError("bf_allpairs_probe: positions are 32-bit (num_bits <= 2^32)")
        row_bytes = self.sliced_row_words(n) * 4
        rows_per_pass = max(32, min((num_bits + 31) // 32 * 32, (max(int(slice_bytes), row_bytes) // row_bytes) // 32 * 32))
        for rb in range(0, num_bits, rows_per_pass):
            re_ = min(num_bits, rb + rows_per_pass)
            sliced = self.bitslice(filters, rb, re_)
            for q0 in range(0, n, batch):
                part = filters[q0: q0 + batch]
                pos, offs = self.extract_positions_packed(part, num_bits)
                hits = self.query_sliced_ranges(sliced, n, rb, re_, pos, [(offs[i], offs[i + 1]) for i in range(len(part))])
                out[q0: q0 + len(part)] += hits.to(torch.int64) & 0xFFFFFFFF
                del pos, hits
            del sliced
        return out

    def bf_rect(self, a: Sequence[torch.Tensor], b: Sequence[torch.Tensor], num_bits: int) -> torch.Tensor:
        """Dense len(a) x len(b) intersection matrix popc(a_i & b_j) (msbt_bf_and_popcount_rect)."""
        out = torch.empty((len(a), len(b)), dtype=torch.int64, device=self.device)
        if len(a) and len(b):
            rc = self._L.msbt_bf_and_popcount_rect(self._h, _ptr_array(a), len(a), _ptr_array(b), len(b), filter_words(num_bits),
                                                   out.data_ptr(), self._stream())
            self._check(rc, "msbt_bf_and_popcount_rect")
        return out

    # ------------------------------------------------------------------ K4
    def extract_positions(self, filt: torch.Tensor, num_bits: int, cap: Optional[int] = None) -> torch.Tensor:
        """Ascending positions of the set bits of one filter (uint32 bit patterns in an int32 tensor).
        Synchronises once to learn the count when `cap` is None."""
        words = filter_words(num_bits)
        count = torch.zeros(1, dtype=torch.int64, device=self.device)
        if cap is None:
            rc = self._L.msbt_bf_extract_positions(self._h, filt.data_ptr(), words, None, 0, count.data_ptr(), self._stream())
            self._check(rc, "msbt_bf_extract_positions")
            cap = int(count.item())
        pos = torch.empty(max(cap, 1), dtype=torch.int32, device=self.device)
        rc = self._L.msbt_bf_extract_positions(self._h, filt.data_ptr(), words, pos.data_ptr(), cap, count.data_ptr(), self._stream())
        self._check(rc, "msbt_bf_extract_positions")
        n = min(int(count.item()), cap)
        return pos[:n]

    def extract_positions_into(self, filters: Sequence[torch.Tensor], num_bits: int, out: torch.Tensor,
                               offsets_out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Ascending set-bit positions of every filter, back to back, written into the caller's int32 tensor `out`
        (entries beyond its size are dropped).  Nothing synchronises: the n + 1 offsets stay on the device (returned)."""
        n = len(filters)
        offs = offsets_out if offsets_out is not None else torch.empty(n + 1, dtype=torch.int64, device=self.device)
        if n == 0:
            return offs
        rc = self._L.msbt_bf_extract_positions_batch(self._h, _ptr_array(filters), n, filter_words(num_bits), out.data_ptr(),
                                                     out.numel(), offs.data_ptr(), self._stream())
        self._check(rc, "msbt_bf_extract_positions_batch")
        return offs

    def extract_positions_packed(self, filters: Sequence[torch.Tensor], num_bits: int, cap: Optional[int] = None
                                 ) -> Tuple[torch.Tensor, List[int]]:
        """Ascending set-bit positions of every filter of a batch, back to back in ONE int32 tensor (uint32 bit
        patterns), plus the n + 1 host offsets that delimit them (three launches for the whole batch).
        `cap`: upper bound on the total count (e.g. the summed k-mer counts); None sizes the buffer with one
        extra counting pass."""
        n = len(filters)
        if n == 0:
            return torch.zeros(1, dtype=torch.int32, device=self.device), [0]
        words = filter_words(num_bits)
        ptrs = _ptr_array(filters)
        offs = torch.empty(n + 1, dtype=torch.int64, device=self.device)
        if cap is None:
            rc = self._L.msbt_bf_extract_positions_batch(self._h, ptrs, n, words, None, 0, offs.data_ptr(), self._stream())
            self._check(rc, "msbt_bf_extract_positions_batch")
            cap = int(offs[n].item())
        pos = torch.empty(max(cap, 1), dtype=torch.int32, device=self.device)
        rc = self._L.msbt_bf_extract_positions_batch(self._h, ptrs, n, words, pos.data_ptr(), cap, offs.data_ptr(), self._stream())
        self._check(rc, "msbt_bf_extract_positions_batch")
        return pos, [min(int(x), cap) for x in offs.cpu().tolist()]

    def extract_positions_batch(self, filters: Sequence[torch.Tensor], num_bits: int, cap: Optional[int] = None) -> List[torch.Tensor]:
        """The same as one int32 view per filter."""
        if len(filters) == 0:
            return []
        pos, h = self.extract_positions_packed(filters, num_bits, cap)
        return [pos[h[i]: h[i + 1]] for i in range(len(filters))]

    def query_batch(self, leaves: Sequence[torch.Tensor], positions: Sequence[torch.Tensor]) -> torch.Tensor:
        """hits[q][l] = number of query q's positions set in leaf l (row-major probe) -> int32[q, l]."""
        nq, nl = len(positions), len(leaves)
        hits = torch.zeros((nq, nl), dtype=torch.int32, device=self.device)
        if nq == 0 or nl == 0:
            return hits
        offs = (ctypes.c_uint64 * (nq + 1))()
        for i, p in enumerate(positions):
            offs[i + 1] = offs[i] + p.numel()
        allpos = torch.cat([p.reshape(-1) for p in positions]) if offs[nq] else torch.zeros(1, dtype=torch.int32, device=self.device)
        rc = self._L.msbt_query_batch(self._h, _ptr_array(leaves), nl, allpos.data_ptr(), offs, nq, hits.data_ptr(), self._stream())
        self._check(rc, "msbt_query_batch")
        return hits

    def query_ranges(self, leaves: Sequence[torch.Tensor], allpos: torch.Tensor, ranges: Sequence[Tuple[int, int]]) -> torch.Tensor:
        """Row-major probe of the queries whose positions are allpos[begin:end] for every (begin, end) of `ranges`
        (any subset, any order, of a packed position batch; nothing is copied) -> int32[q, l]."""
        nq, nl = len(ranges), len(leaves)
        hits = torch.zeros((nq, nl), dtype=torch.int32, device=self.device)
        if nq == 0 or nl == 0:
            return hits
        rng = (ctypes.c_uint64 * (2 * nq))()
        for i, (b, e) in enumerate(ranges):
            rng[2 * i], rng[2 * i + 1] = int(b), int(e)
        rc = self._L.msbt_query_batch_ranges(self._h, _ptr_array(leaves), nl, allpos.data_ptr(), rng, nq, hits.data_ptr(), self._stream())
        self._check(rc, "msbt_query_batch_ranges")
        return hits

    def query_sliced_ranges(self, sliced: torch.Tensor, n_leaves: int, row_begin: int, row_end: int,
                            allpos: torch.Tensor, ranges: Sequence[Tuple[int, int]]) -> torch.Tensor:
        """Bit-sliced probe with an explicit (begin, end) range of `allpos` per query -> int32[q, l]."""
        nq = len(ranges)
        hits = torch.empty((nq, n_leaves), dtype=torch.int32, device=self.device)
        if nq == 0 or n_leaves == 0:
            return hits
        out = []
        for q0 in range(0, nq, 65535):  # one launch per 65,535 queries (grid limit)
            part = ranges[q0: q0 + 65535]
            rng = (ctypes.c_uint64 * (2 * len(part)))()
            for i, (b, e) in enumerate(part):
                rng[2 * i], rng[2 * i + 1] = int(b), int(e)
            rc = self._L.msbt_query_batch_sliced_ranges(self._h, sliced.data_ptr(), n_leaves, row_begin, row_end, allpos.data_ptr(),
                                                        rng, len(part), hits[q0: q0 + len(part)].data_ptr(), self._stream())
            self._check(rc, "msbt_query_batch_sliced_ranges")
        return hits

    def sliced_row_words(self, n_leaves: int) -> int:
        """u32 words per row of the bit-sliced matrix for `n_leaves` leaves (ceil(L/32) rounded up to 8 words)."""
        return int(self._L.msbt_sliced_row_words(n_leaves))

    def bitslice(self, leaves: Sequence[torch.Tensor], row_begin: int, row_end: int) -> torch.Tensor:
        """Bit-sliced leaf matrix for rows [row_begin, row_end): int32[rows, sliced_row_words(L)] (rows padded to 32 bytes)."""
        lw = self.sliced_row_words(len(leaves))
        out = torch.empty((max(row_end - row_begin, 0), lw), dtype=torch.int32, device=self.device)
        rc = self._L.msbt_bitslice_build(self._h, _ptr_array(leaves), len(leaves), row_begin, row_end, out.data_ptr(), self._stream())
        self._check(rc, "msbt_bitslice_build")
        return out

    def query_batch_sliced(self, sliced: torch.Tensor, n_leaves: int, row_begin: int, row_end: int,
                           positions: Sequence[torch.Tensor]) -> torch.Tensor:
        """Partial hits of the positions that fall into rows [row_begin, row_end) -> int32[q, l]."""
        offs = [0]
        for p in positions:
            offs.append(offs[-1] + p.numel())
        allpos = torch.cat([p.reshape(-1) for p in positions]) if offs[-1] else torch.zeros(1, dtype=torch.int32, device=self.device)
        return self.query_sliced_packed(sliced, n_leaves, row_begin, row_end, allpos, offs)

    def query_sliced_packed(self, sliced: torch.Tensor, n_leaves: int, row_begin: int, row_end: int,
                            allpos: torch.Tensor, offsets: Sequence[int]) -> torch.Tensor:
        """query_batch_sliced on positions that already sit back to back in one tensor (extract_positions_packed)."""
        nq = len(offsets) - 1
        hits = torch.zeros((max(nq, 0), n_leaves), dtype=torch.int32, device=self.device)
        if nq <= 0 or n_leaves == 0:
            return hits
        offs = (ctypes.c_uint64 * (nq + 1))(*[int(x) for x in offsets])
        rc = self._L.msbt_query_batch_sliced(self._h, sliced.data_ptr(), n_leaves, row_begin, row_end, allpos.data_ptr(),
                                             offs, nq, hits.data_ptr(), self._stream())
        self._check(rc, "msbt_query_batch_sliced")
        return hits


_DEFAULT: dict = {}


def default_context(device: Optional[int] = None) -> Context:
    """Process-wide Context per device (created on first use)."""
    if not torch.cuda.is_available():
        raise RuntimeError("metasbt_b200 needs a CUDA device: the hot path has no CPU fallback")
    idx = torch.cuda.current_device() if device is None else device
    if idx not in _DEFAULT:
        _DEFAULT[idx] = Context(idx)
    return _DEFAULT[idx]
