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

