// tests/emu/emu_kmer.cpp -- TEST INFRASTRUCTURE.  Compiles the device-side per-thread k-mer
// code (metasbt_b200/csrc/msbt_kmer.h, msbt_spec.h) for the host and runs it in a plain loop
// over "threads", so that the walker / block / validity-mask / hash / position logic can be
// checked against the oracle in the CPU test suite (no GPU in the dev container).
#include <stdint.h>
#include <string.h>
#define __restrict__
#include "../../metasbt_b200/csrc/msbt_kmer.h"

template <int KW>
static void run_walker(const uint64_t *packed, const uint32_t *ends, uint32_t n_runs, uint32_t n_bases, uint32_t k,
                       const msbt_pos_spec &ps, uint32_t *out32) {
    const uint32_t n_words = (n_bases + 31) / 32;
    for (uint32_t w = 0; w < n_words; w++) {
        const uint32_t s0 = w * 32;
        const uint32_t vm = valid_mask(ends, n_runs, s0, k);
        if (!vm) continue;
        KmerWalker<KW> kw;
        kw.init(packed + w, k);
        for (uint32_t j = 0; j < 32; j++) {
            if ((vm >> j) & 1) {
                uint64_t lo, hi;
                kw.canonical(lo, hi);
                uint64_t pos = msbt_position(msbt_kmer_hash(lo, hi, k, ps.seed), ps);
                if (pos < ps.num_bits) out32[pos >> 5] |= 1u << (pos & 31);
            }
            kw.roll(j);
        }
    }
}

template <int KW, int HV, bool POW2>
struct BlockRunner {
    const msbt_pos_spec &ps;
    uint32_t k, len, vm;
    uint32_t *out32;
    template <int J>
    void one(const KmerBlock<KW> &kb, uint32_t g) {
        if ((vm >> (8 * g + J)) & 1) {
            uint64_t lo, hi;
            kb.template get<J>(lo, hi);
            uint64_t pos = msbt_position_v<POW2>(msbt_kmer_hash_v<HV>(lo, hi, len, ps.seed), ps);
            if (pos < ps.num_bits) out32[pos >> 5] |= 1u << (pos & 31);
        }
    }
};

template <int KW, int HV, bool POW2>
static void run_block(const uint64_t *packed, const uint32_t *ends, uint32_t n_runs, uint32_t n_bases, uint32_t k,
                      const msbt_pos_spec &ps, uint32_t *out32) {
    const uint32_t n_words = (n_bases + 31) / 32;
    for (uint32_t w = 0; w < n_words; w++) {
        const uint32_t vm = valid_mask(ends, n_runs, w * 32, k);
        if (!vm) continue;
        KmerBlock<KW> kb;
        kb.init(packed + w, k);
        BlockRunner<KW, HV, POW2> br{ps, k, (k + 3) / 4, vm, out32};
        for (uint32_t g = 0; g < 4; g++) {
            br.template one<0>(kb, g); br.template one<1>(kb, g); br.template one<2>(kb, g); br.template one<3>(kb, g);
            br.template one<4>(kb, g); br.template one<5>(kb, g); br.template one<6>(kb, g); br.template one<7>(kb, g);
            kb.advance8();
        }
    }
}

extern "C" int emu_makebf(const uint64_t *packed, const uint32_t *ends, uint32_t n_runs, uint32_t n_bases, uint32_t k,
                          uint64_t seed, uint64_t modulus, uint64_t num_bits, uint64_t *words, int impl) {
    if (k < 1 || k > 64) return -1;
    memset(words, 0, ((num_bits + 63) / 64) * 8);
    msbt_pos_spec ps = msbt_make_pos_spec(modulus, num_bits, seed);
    uint32_t *o = (uint32_t *)words;
    if (impl == 0) {
        if (k <= 32) run_walker<1>(packed, ends, n_runs, n_bases, k, ps, o);
        else run_walker<2>(packed, ends, n_runs, n_bases, k, ps, o);
        return 0;
    }
    const int hv = msbt_hash_variant(k);
#define RUN(KW, HV) (ps.pow2 ? run_block<KW, HV, true>(packed, ends, n_runs, n_bases, k, ps, o) : run_block<KW, HV, false>(packed, ends, n_runs, n_bases, k, ps, o))
    if (hv == 0) RUN(1, 0);
    else if (hv == 1) RUN(2, 1);
    else RUN(2, 2);
#undef RUN
    return 0;
}

// The generic K1 kernel's per-thread code (run-time msbt_hash_spec switches), for the host.
template <int KW>
static void run_generic(const uint64_t *packed, const uint32_t *ends, uint32_t n_runs, uint32_t n_bases, uint32_t k,
                        const msbt_pos_spec &ps, uint32_t key_len, uint32_t second_half, uint32_t forward_only, uint32_t *out32) {
    const uint32_t n_words = (n_bases + 31) / 32;
    for (uint32_t w = 0; w < n_words; w++) {
        const uint32_t vm = valid_mask(ends, n_runs, w * 32, k);
        if (!vm) continue;
        KmerWalker<KW> kw;
        kw.init(packed + w, k);
        for (uint32_t j = 0; j < 32; j++) {
            if ((vm >> j) & 1) {
                uint64_t lo, hi;
                if (forward_only) { lo = kw.fl; hi = kw.fh; }
                else kw.canonical(lo, hi);
                uint64_t pos = msbt_position(msbt_kmer_hash_generic(lo, hi, key_len, ps.seed, second_half), ps);
                if (pos < ps.num_bits) out32[pos >> 5] |= 1u << (pos & 31);
            }
            kw.roll(j);
        }
    }
}

extern "C" int emu_makebf_generic(const uint64_t *packed, const uint32_t *ends, uint32_t n_runs, uint32_t n_bases, uint32_t k,
                                  uint64_t seed, uint64_t modulus, uint64_t num_bits, uint64_t *words,
                                  uint32_t key_rule, uint32_t second_half, uint32_t forward_only) {
    if (k < 1 || k > 64) return -1;
    memset(words, 0, ((num_bits + 63) / 64) * 8);
    msbt_pos_spec ps = msbt_make_pos_spec(modulus, num_bits, seed);
    const uint32_t key_len = key_rule ? (k <= 32 ? 8u : 16u) : (k + 3u) / 4u;
    if (k <= 32) run_generic<1>(packed, ends, n_runs, n_bases, k, ps, key_len, second_half, forward_only, (uint32_t *)words);
    else run_generic<2>(packed, ends, n_runs, n_bases, k, ps, key_len, second_half, forward_only, (uint32_t *)words);
    return 0;
}
