// msbt_spec.h -- the ONE place that fixes every HowDeSBT-side convention of the hot path
// (SURVEY.md Appendix A): base codes, canonical rule, hash function, seed use, modulus rule,
// bit order inside filter words and the `.bf` header constants.  Host and device code both
// include it; nothing else in metasbt_b200/ re-states any of these.  If a real `howdesbt`
// ever disagrees with an item here, this header is the only file that has to change
// (the CPU checker in oracle/ is a separate, independent restatement of the same spec).
//
// Reference call site that fixes the parameters: /root/reference/metasbt/core.py:3920-3931
//   howdesbt makebf --k=K --min=C --bits=M --hashes=1 --seed=0,0
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define MSBT_HD __host__ __device__ __forceinline__
#define MSBT_HDM __host__ __device__ __forceinline__  // member functions
#else
#define MSBT_HD static inline
#define MSBT_HDM inline
#endif

// ---- base codes (A1): A=0 C=1 G=2 T=3, case-insensitive; everything else breaks a run.
#define MSBT_BASE_INVALID 4

// ---- packed genome layout (ours; the reference hands howdesbt a FASTA path instead):
// valid bases only, 2 bits each, 32 per u64 word, FIRST base in the TOP two bits, so that
// the k-mer starting at a word boundary is simply `word >> (64 - 2k)` with the first base
// most significant -- the same integer the canonical comparison and the hash are defined on.
#define MSBT_BASES_PER_WORD 32
#define MSBT_PACK_PAD_WORDS 3  // zero words after every genome so w+1, w+2 reads stay in bounds

// ---- `.bf` file constants (A2); see metasbt_b200/bffile.py for the reader/writer.
#define MSBT_BF_MAGIC 0xD532006662544253ULL
#define MSBT_BF_HEADER_SIZE 0x70u
#define MSBT_BF_VERSION 2u
#define MSBT_BF_KIND_SIMPLE 1u
#define MSBT_BF_COMP_UNCOMPRESSED 1u

MSBT_HD uint64_t msbt_rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }

MSBT_HD uint64_t msbt_fmix64(uint64_t k) {
    k ^= k >> 33;
    k *= 0xff51afd7ed558ccdULL;
    k ^= k >> 33;
    k *= 0xc4ceb9fe1a85ec53ULL;
    k ^= k >> 33;
    return k;
}

// Reverse the order of the 32 two-bit groups of x and complement each base (c -> 3-c).
MSBT_HD uint64_t msbt_revcomp_word(uint64_t x) {
    x = ~x;
    x = ((x >> 2) & 0x3333333333333333ULL) | ((x & 0x3333333333333333ULL) << 2);
    x = ((x >> 4) & 0x0F0F0F0F0F0F0F0FULL) | ((x & 0x0F0F0F0F0F0F0F0FULL) << 4);
    x = ((x >> 8) & 0x00FF00FF00FF00FFULL) | ((x & 0x00FF00FF00FF00FFULL) << 8);
    x = ((x >> 16) & 0x0000FFFF0000FFFFULL) | ((x & 0x0000FFFF0000FFFFULL) << 16);
    return (x >> 32) | (x << 32);
}

// ---- hash (A1): first 64-bit half of MurmurHash3_x64_128(seed) over the canonical k-mer's
// 2-bit packed little-endian bytes (last base in the two lowest bits of byte 0), length
// (k+3)/4 bytes.  `lo`/`hi` are the low/high 64 bits of the k-mer integer (hi == 0 for k<=32).
MSBT_HD uint64_t msbt_kmer_hash(uint64_t lo, uint64_t hi, uint32_t k, uint32_t seed) {
    const uint64_t c1 = 0x87c37b91114253d5ULL, c2 = 0x4cf5ad432745937fULL;
    const uint32_t len = (k + 3u) >> 2;
    uint64_t h1 = seed, h2 = seed;
    uint64_t k1 = lo, k2 = hi;
    k1 *= c1; k1 = msbt_rotl64(k1, 31); k1 *= c2;
    if (len > 8) { k2 *= c2; k2 = msbt_rotl64(k2, 33); k2 *= c1; }
    if (len == 16) {  // one full 16-byte block, empty tail
        h1 ^= k1; h1 = msbt_rotl64(h1, 27); h1 += h2; h1 = h1 * 5 + 0x52dce729;
        h2 ^= k2; h2 = msbt_rotl64(h2, 31); h2 += h1; h2 = h2 * 5 + 0x38495ab5;
    } else {          // tail only
        if (len > 8) h2 ^= k2;
        h1 ^= k1;
    }
    h1 ^= len; h2 ^= len;
    h1 += h2; h2 += h1;
    h1 = msbt_fmix64(h1); h2 = msbt_fmix64(h2);
    return h1 + h2;
}

// The same digest with every HowDeSBT-side choice that is NOT pinned against a real binary as a run-time argument
// (include/msbt.h: msbt_hash_spec): `len` = number of key bytes hashed (1..16), `second_half` = which 64-bit half of the
// 128-bit digest is the hash value.  msbt_kmer_hash(lo, hi, k, seed) == msbt_kmer_hash_generic(lo, hi, (k+3)/4, seed, 0).
// Used by the generic K1 kernel that serves contexts created with a non-default spec; the tuned kernels keep the
// default compiled in.
MSBT_HD uint64_t msbt_kmer_hash_generic(uint64_t lo, uint64_t hi, uint32_t len, uint32_t seed, uint32_t second_half) {
    const uint64_t c1 = 0x87c37b91114253d5ULL, c2 = 0x4cf5ad432745937fULL;
    uint64_t h1 = seed, h2 = seed;
    uint64_t k1 = lo, k2 = hi;
    // bytes beyond `len` do not belong to the key
    if (len < 8) k1 &= (1ULL << (8 * len)) - 1;
    if (len <= 8) k2 = 0;
    else if (len < 16) k2 &= (1ULL << (8 * (len - 8))) - 1;
    k1 *= c1; k1 = msbt_rotl64(k1, 31); k1 *= c2;
    if (len > 8) { k2 *= c2; k2 = msbt_rotl64(k2, 33); k2 *= c1; }
    if (len == 16) {
        h1 ^= k1; h1 = msbt_rotl64(h1, 27); h1 += h2; h1 = h1 * 5 + 0x52dce729;
        h2 ^= k2; h2 = msbt_rotl64(h2, 31); h2 += h1; h2 = h2 * 5 + 0x38495ab5;
    } else {
        if (len > 8) h2 ^= k2;
        if (len > 0) h1 ^= k1;
    }
    h1 ^= len; h2 ^= len;
    h1 += h2; h2 += h1;
    h1 = msbt_fmix64(h1); h2 = msbt_fmix64(h2);
    h1 += h2; h2 += h1;
    return second_half ? h2 : h1;
}

// ---- position (A1): pos = h % modulus; the bit is set/probed only if pos < num_bits.
// `magic` = floor((2^64-1)/modulus) lets the device replace the 64-bit division by one
// multiply-high and one correction (exact, see msbt_mod_magic); power-of-two moduli mask.
typedef struct {
    uint64_t modulus;
    uint64_t magic;
    uint64_t num_bits;
    uint32_t pow2;  // 1 if modulus is a power of two
    uint32_t seed;  // MurmurHash3 takes a 32-bit seed; hashSeed1 is truncated to it
} msbt_pos_spec;

static inline msbt_pos_spec msbt_make_pos_spec(uint64_t modulus, uint64_t num_bits, uint64_t seed) {
    msbt_pos_spec s;
    s.modulus = modulus;
    s.num_bits = num_bits;
    s.pow2 = (modulus & (modulus - 1)) == 0;
    s.magic = s.pow2 ? 0 : (~0ULL) / modulus;
    s.seed = (uint32_t)seed;
    return s;
}

MSBT_HD uint64_t msbt_mulhi64(uint64_t a, uint64_t b) {
#if defined(__CUDA_ARCH__)
    return __umul64hi(a, b);
#else
    return (uint64_t)(((unsigned __int128)a * b) >> 64);
#endif
}

MSBT_HD uint64_t msbt_position(uint64_t h, const msbt_pos_spec& s) {
    if (s.pow2) return h & (s.modulus - 1);
    // q in {Q-1, Q} for Q = floor(h / modulus)  =>  r in [0, 2*modulus)
    uint64_t q = msbt_mulhi64(h, s.magic);
    uint64_t r = h - q * s.modulus;
    return r >= s.modulus ? r - s.modulus : r;
}
