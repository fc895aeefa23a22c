// msbt_kmer.h -- per-thread k-mer machinery shared by every K1/K4 kernel: the rolling
// canonical k-mer walker over the packed genome layout of msbt_spec.h and the run-table
// validity mask.  Everything is MSBT_HD so that tests/emu/ can compile the very same code
// for the host and check it against the oracle without a GPU.
#pragma once
#include "msbt_spec.h"

// 64-bit funnel: bits of (hi:lo) << s, upper 64 bits; s in [0, 64]
MSBT_HD uint64_t funnel_hi(uint64_t hi, uint64_t lo, uint32_t s) {
    if (s == 0) return hi;
    if (s >= 64) return lo;
    return (hi << s) | (lo >> (64 - s));
}

// Per-thread state for the 32 k-mer starts that begin inside packed word `w` of a genome.
// KW = 1: k <= 32 (one 64-bit word per k-mer); KW = 2: 32 < k <= 64.
template <int KW>
struct KmerWalker {
    uint64_t fl, fh, rl, rh;  // forward / reverse-complement k-mer, low and high words
    uint64_t stream;          // the 32 bases that enter the window for j = 1..32
    uint64_t mask_l, mask_h;
    uint32_t rc_shift;        // position of the base entering the reverse complement

    MSBT_HDM void init(const uint64_t *__restrict__ p, uint32_t k) {
        uint64_t w0 = p[0], w1 = p[1];
        if (KW == 1) {
            fl = w0 >> (64 - 2 * k);
            fh = 0;
            rl = msbt_revcomp_word(fl) >> (64 - 2 * k);
            rh = 0;
            stream = funnel_hi(w0, w1, 2 * k);
            mask_l = (k == 32) ? ~0ULL : ((1ULL << (2 * k)) - 1);
            mask_h = 0;
            rc_shift = 2 * k - 2;
        } else {
            uint64_t w2 = p[2];
            uint32_t s = 128 - 2 * k;  // in [0, 62]
            // forward = (w0:w1) >> s
            fh = s ? (w0 >> s) : w0;
            fl = s ? ((w1 >> s) | (w0 << (64 - s))) : w1;
            // reverse complement of the 128-bit value, then >> s
            uint64_t xh = msbt_revcomp_word(fl), xl = msbt_revcomp_word(fh);
            rh = s ? (xh >> s) : xh;
            rl = s ? ((xl >> s) | (xh << (64 - s))) : xl;
            stream = funnel_hi(w1, w2, 2 * k - 64);
            mask_l = ~0ULL;
            mask_h = (k == 64) ? ~0ULL : ((1ULL << (2 * k - 64)) - 1);
            rc_shift = 2 * k - 2 - 64;
        }
    }

    MSBT_HDM void canonical(uint64_t &lo, uint64_t &hi) const {
        if (KW == 1) {
            lo = fl < rl ? fl : rl;
            hi = 0;
        } else {
            bool fwd_less = fh < rh || (fh == rh && fl < rl);
            lo = fwd_less ? fl : rl;
            hi = fwd_less ? fh : rh;
        }
    }

    // slide the window by one base; j = index of the k-mer just consumed (0..31)
    MSBT_HDM void roll(uint32_t j) {
        uint64_t b = (stream >> (62 - 2 * j)) & 3;
        if (KW == 1) {
            fl = ((fl << 2) | b) & mask_l;
            rl = (rl >> 2) | ((b ^ 3) << rc_shift);
        } else {
            fh = ((fh << 2) | (fl >> 62)) & mask_h;
            fl = (fl << 2) | b;
            rl = (rl >> 2) | (rh << 62);
            rh = (rh >> 2) | ((b ^ 3) << rc_shift);
        }
    }
};

// Which of the 32 k-mer starts s0 .. s0+31 lie fully inside a run and inside the genome.
MSBT_HD uint32_t valid_mask(const uint32_t *__restrict__ ends, uint32_t n_runs,
                                               uint32_t s0, uint32_t k) {
    // first run whose end is > s0
    uint32_t lo = 0, hi = n_runs;
    while (lo < hi) {
        uint32_t mid = (lo + hi) >> 1;
        if (ends[mid] > s0) hi = mid; else lo = mid + 1;
    }
    if (lo >= n_runs) return 0;
    uint32_t r = lo, limit = ends[r];
    if ((uint64_t)s0 + 31 + k <= limit) return 0xFFFFFFFFu;
    uint32_t m = 0;
    for (uint32_t j = 0; j < 32; j++) {
        uint32_t i = s0 + j;
        while (r < n_runs && i >= limit) {
            r++;
            if (r < n_runs) limit = ends[r];
        }
        if (r >= n_runs) break;
        if ((uint64_t)i + k <= limit) m |= 1u << j;
    }
    return m;
}


// ---------------------------------------------------------------------------------------------
// KmerBlock: the same 32 k-mers per packed word as KmerWalker, but with no loop-carried rolling
// state: every k-mer is cut out of two pre-aligned streams with compile-time funnel shifts
//   T = forward stream >> (64*KW - 2k)      fwd_j = top words of (T << 2j), masked to 2k bits
//   R = reverse-complement stream (base i at bit 2i)   rc_j = low words of (R >> 2j), masked
// so that, once unrolled, a k-mer costs 2 funnels + masks + one compare instead of two dependent
// shift/or chains with run-time shift amounts.  Processed in 4 groups of 8: get<J>() for J = 0..7,
// then advance8().
MSBT_HD uint64_t funnel_lo(uint64_t hi, uint64_t lo, uint32_t s) {  // low 64 bits of (hi:lo) >> s, s in [0, 63]
    return s ? ((lo >> s) | (hi << (64 - s))) : lo;
}

template <int KW>
struct KmerBlock {
    uint64_t t0, t1, t2;  // T, most significant word last used: KW=1 -> (t1:t0), KW=2 -> (t2:t1:t0)
    uint64_t r0, r1, r2;  // R, r0 least significant
    uint64_t mask;        // KW=1: mask of the k-mer word; KW=2: mask of the high word

    MSBT_HDM void init(const uint64_t *__restrict__ p, uint32_t k) {
        const uint64_t w0 = p[0], w1 = p[1];
        if (KW == 1) {
            const uint32_t s = 64 - 2 * k;  // [0, 62]
            t1 = s ? (w0 >> s) : w0;
            t0 = s ? ((w1 >> s) | (w0 << (64 - s))) : w1;
            t2 = 0;
            r0 = msbt_revcomp_word(w0);
            r1 = msbt_revcomp_word(w1);
            r2 = 0;
            mask = (k == 32) ? ~0ULL : ((1ULL << (2 * k)) - 1);
        } else {
            const uint64_t w2 = p[2];
            const uint32_t s = 128 - 2 * k;  // [0, 62]
            t2 = s ? (w0 >> s) : w0;
            t1 = s ? ((w1 >> s) | (w0 << (64 - s))) : w1;
            t0 = s ? ((w2 >> s) | (w1 << (64 - s))) : w2;
            r0 = msbt_revcomp_word(w0);
            r1 = msbt_revcomp_word(w1);
            r2 = msbt_revcomp_word(w2);
            mask = (k == 64) ? ~0ULL : ((1ULL << (2 * k - 64)) - 1);
        }
    }

    // canonical k-mer number J (0..7) of the current group
    template <int J>
    MSBT_HDM void get(uint64_t &lo, uint64_t &hi) const {
        constexpr uint32_t c = 2 * J;
        if (KW == 1) {
            const uint64_t f = funnel_hi(t1, t0, c) & mask;
            const uint64_t r = funnel_lo(r1, r0, c) & mask;
            lo = f < r ? f : r;
            hi = 0;
        } else {
            const uint64_t fh = funnel_hi(t2, t1, c) & mask, fl = funnel_hi(t1, t0, c);
            const uint64_t rh = funnel_lo(r2, r1, c) & mask, rl = funnel_lo(r1, r0, c);
            const bool fwd_less = fh < rh || (fh == rh && fl < rl);
            lo = fwd_less ? fl : rl;
            hi = fwd_less ? fh : rh;
        }
    }

    MSBT_HDM void advance8() {
        if (KW == 1) {
            t1 = (t1 << 16) | (t0 >> 48);
            t0 <<= 16;
            r0 = (r0 >> 16) | (r1 << 48);
            r1 >>= 16;
        } else {
            t2 = (t2 << 16) | (t1 >> 48);
            t1 = (t1 << 16) | (t0 >> 48);
            t0 <<= 16;
            r0 = (r0 >> 16) | (r1 << 48);
            r1 = (r1 >> 16) | (r2 << 48);
            r2 >>= 16;
        }
    }
};

// Hash specialised on the length class of the packed k-mer: HV = 0: <= 8 bytes (k <= 32),
// 1: 9..15 bytes (33 <= k <= 60), 2: 16 bytes (61 <= k <= 64).  Same function as msbt_kmer_hash.
template <int HV>
MSBT_HD uint64_t msbt_kmer_hash_v(uint64_t lo, uint64_t hi, uint32_t len, uint32_t seed) {
    const uint64_t c1 = 0x87c37b91114253d5ULL, c2 = 0x4cf5ad432745937fULL;
    uint64_t h1 = seed, h2 = seed;
    uint64_t k1 = lo * c1;
    k1 = msbt_rotl64(k1, 31);
    k1 *= c2;
    if (HV >= 1) {
        uint64_t k2 = hi * c2;
        k2 = msbt_rotl64(k2, 33);
        k2 *= c1;
        if (HV == 2) {
            h1 ^= k1; h1 = msbt_rotl64(h1, 27); h1 += h2; h1 = h1 * 5 + 0x52dce729;
            h2 ^= k2; h2 = msbt_rotl64(h2, 31); h2 += h1; h2 = h2 * 5 + 0x38495ab5;
        } else {
            h2 ^= k2;
            h1 ^= k1;
        }
    } else {
        h1 ^= k1;
    }
    h1 ^= len; h2 ^= len;
    h1 += h2; h2 += h1;
    h1 = msbt_fmix64(h1); h2 = msbt_fmix64(h2);
    return h1 + h2;
}

MSBT_HD int msbt_hash_variant(uint32_t k) { return k <= 32 ? 0 : (k <= 60 ? 1 : 2); }

template <bool POW2>
MSBT_HD uint64_t msbt_position_v(uint64_t h, const msbt_pos_spec &s) {
    if (POW2) return h & (s.modulus - 1);
    uint64_t q = msbt_mulhi64(h, s.magic);
    uint64_t r = h - q * s.modulus;
    return r >= s.modulus ? r - s.modulus : r;
}
