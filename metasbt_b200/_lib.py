"""ctypes binding of libmsbt.so (include/msbt.h).  No fallback: if the shared library is
missing the import of any hot-path op fails loudly with the build command to run."""

from __future__ import annotations

import ctypes
import os
import subprocess
import sys
from typing import List

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG)
SO_PATH = os.environ.get("MSBT_LIB") or os.path.join(_PKG, "libmsbt.so")  # MSBT_LIB: kernel-tuning builds only
SOURCES = [os.path.join(_PKG, "csrc", "msbt_kernels.cu")]
import glob as _glob
# every header the translation unit includes (a stale libmsbt.so must never survive an edit of a kernel header)
HEADERS = sorted(_glob.glob(os.path.join(_PKG, "csrc", "*.cuh")) + _glob.glob(os.path.join(_PKG, "csrc", "*.h"))) \
    + [os.path.join(_ROOT, "include", "msbt.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]

# every symbol include/msbt.h declares; tests check that the library exports all of them
SYMBOLS = [
    "msbt_abi_version", "msbt_hash_spec_default", "msbt_ctx_create", "msbt_ctx_get_spec", "msbt_ctx_destroy", "msbt_last_error", "msbt_sync",
    "msbt_launch_count", "msbt_fasta_pack", "msbt_fasta_pack_device", "msbt_kmer_bloom_build", "msbt_bf_or_reduce",
    "msbt_bf_popcount", "msbt_bf_occurrence", "msbt_bf_and_popcount_1vN", "msbt_bf_and_popcount_pairs", "msbt_bf_and_popcount_allpairs", "msbt_bf_and_popcount_rect",
    "msbt_query_batch_ranges", "msbt_query_batch_sliced_ranges", "msbt_positions_expand",
    "msbt_bf_extract_positions", "msbt_bf_extract_positions_batch", "msbt_query_batch", "msbt_bitslice_build", "msbt_query_batch_sliced", "msbt_sliced_row_words",
]

MSBT_OK, MSBT_EINVAL, MSBT_ENOMEM, MSBT_ENOSPC, MSBT_ECUDA = 0, -22, -12, -28, -5
PACK_PAD_WORDS = 3


class GenomeDesc(ctypes.Structure):
    """msbt_genome_desc"""
    _fields_ = [
        ("packed_word_off", ctypes.c_uint64),
        ("n_bases", ctypes.c_uint64),
        ("run_off", ctypes.c_uint64),
        ("n_runs", ctypes.c_uint32),
        ("filter_index", ctypes.c_uint32),
    ]


class HashSpecStruct(ctypes.Structure):
    """msbt_hash_spec"""
    _fields_ = [
        ("struct_size", ctypes.c_uint32),
        ("key_bytes", ctypes.c_uint32),
        ("digest_half", ctypes.c_uint32),
        ("canonical", ctypes.c_uint32),
        ("bf_magic", ctypes.c_uint64),
        ("bf_version", ctypes.c_uint32),
        ("bf_header_size", ctypes.c_uint32),
    ]


def needs_build() -> bool:
    if not os.path.isfile(SO_PATH):
        return True
    t = os.path.getmtime(SO_PATH)
    return any(os.path.isfile(p) and os.path.getmtime(p) > t for p in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False) -> str:
    """nvcc -gencode arch=compute_100a,code=sm_100a ... -> metasbt_b200/libmsbt.so (in-tree)."""
    if force or needs_build():
        nvcc = os.environ.get("NVCC", "nvcc")
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", SO_PATH] + SOURCES
        subprocess.check_call(cmd)
    return SO_PATH


_LIB = None


def lib() -> ctypes.CDLL:
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.isfile(SO_PATH):
        raise RuntimeError(
            f"{SO_PATH} is missing: the CUDA extension is the only implementation of the hot path "
            "(there is no CPU fallback). Build it with `python -c 'import __graft_entry__ as g; g.build()'`."
        )
    L = ctypes.CDLL(SO_PATH)
    u64, u32, i32 = ctypes.c_uint64, ctypes.c_uint32, ctypes.c_int
    vp, sz = ctypes.c_void_p, ctypes.c_size_t
    P = ctypes.POINTER
    L.msbt_abi_version.argtypes = []
    L.msbt_abi_version.restype = i32
    L.msbt_hash_spec_default.argtypes = [P(HashSpecStruct)]
    L.msbt_hash_spec_default.restype = None
    L.msbt_ctx_create.argtypes = [i32, P(HashSpecStruct), P(vp)]
    L.msbt_ctx_create.restype = i32
    L.msbt_ctx_get_spec.argtypes = [vp, P(HashSpecStruct)]
    L.msbt_ctx_get_spec.restype = i32
    L.msbt_ctx_destroy.argtypes = [vp]
    L.msbt_ctx_destroy.restype = None
    L.msbt_last_error.argtypes = [vp]
    L.msbt_last_error.restype = ctypes.c_char_p
    L.msbt_sync.argtypes = [vp, vp]
    L.msbt_sync.restype = i32
    L.msbt_launch_count.argtypes = [vp]
    L.msbt_launch_count.restype = u64
    L.msbt_fasta_pack.argtypes = [ctypes.c_char_p, sz, u32, vp, sz, vp, sz, P(u64), P(u64)]
    L.msbt_fasta_pack.restype = i32
    L.msbt_fasta_pack_device.argtypes = [vp, vp, P(u64), u32, u32, vp, u64, vp, u64, P(GenomeDesc), vp]
    L.msbt_fasta_pack_device.restype = i32
    L.msbt_kmer_bloom_build.argtypes = [vp, vp, vp, P(GenomeDesc), u32, u32, u64, u64, u64, u32, vp, u64, i32, vp]
    L.msbt_kmer_bloom_build.restype = i32
    L.msbt_bf_or_reduce.argtypes = [vp, P(vp), u32, u64, vp, vp]
    L.msbt_bf_or_reduce.restype = i32
    L.msbt_bf_popcount.argtypes = [vp, P(vp), u32, u64, vp, vp]
    L.msbt_bf_popcount.restype = i32
    L.msbt_bf_occurrence.argtypes = [vp, P(vp), u32, u64, vp, vp]
    L.msbt_bf_occurrence.restype = i32
    L.msbt_bf_and_popcount_1vN.argtypes = [vp, vp, P(vp), u32, u64, vp, vp, vp]
    L.msbt_bf_and_popcount_1vN.restype = i32
    L.msbt_bf_and_popcount_pairs.argtypes = [vp, P(vp), P(vp), u32, u64, vp, vp, vp]
    L.msbt_bf_and_popcount_pairs.restype = i32
    L.msbt_bf_and_popcount_rect.argtypes = [vp, P(vp), u32, P(vp), u32, u64, vp, vp]
    L.msbt_bf_and_popcount_rect.restype = i32
    L.msbt_bf_and_popcount_allpairs.argtypes = [vp, P(vp), u32, u64, vp, vp]
    L.msbt_bf_and_popcount_allpairs.restype = i32
    L.msbt_bf_extract_positions.argtypes = [vp, vp, u64, vp, u64, vp, vp]
    L.msbt_bf_extract_positions.restype = i32
    L.msbt_bf_extract_positions_batch.argtypes = [vp, P(vp), u32, u64, vp, u64, vp, vp]
    L.msbt_bf_extract_positions_batch.restype = i32
    L.msbt_query_batch.argtypes = [vp, P(vp), u32, vp, P(u64), u32, vp, vp]
    L.msbt_query_batch.restype = i32
    L.msbt_bitslice_build.argtypes = [vp, P(vp), u32, u64, u64, vp, vp]
    L.msbt_bitslice_build.restype = i32
    L.msbt_query_batch_sliced.argtypes = [vp, vp, u32, u64, u64, vp, P(u64), u32, vp, vp]
    L.msbt_query_batch_sliced.restype = i32
    L.msbt_query_batch_ranges.argtypes = [vp, P(vp), u32, vp, P(u64), u32, vp, vp]
    L.msbt_query_batch_ranges.restype = i32
    L.msbt_query_batch_sliced_ranges.argtypes = [vp, vp, u32, u64, u64, vp, P(u64), u32, vp, vp]
    L.msbt_query_batch_sliced_ranges.restype = i32
    L.msbt_positions_expand.argtypes = [vp, u64, u64, vp]
    L.msbt_positions_expand.restype = i32
    L.msbt_sliced_row_words.argtypes = [u32]
    L.msbt_sliced_row_words.restype = u32
    _LIB = L
    return L


def loaded_path() -> str:
    """Path of the shared object actually mapped into this process (for the bench report)."""
    lib()
    return SO_PATH


def exported_symbols() -> List[str]:
    L = ctypes.CDLL(SO_PATH)
    return [s for s in SYMBOLS if hasattr(L, s)]


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
    print(SO_PATH)
