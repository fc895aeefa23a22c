// tests/emu/emu_kmer.cpp -- TEST INFRASTRUCTURE.  Compiles the device-side per-thread k-mer
// code (metasbt_b200/csrc/msbt_kmer.h, msbt_spec.h) for the host and runs it in a plain loop
// over "threads", so that the walker / validity-mask / hash / position logic can be checked
// against the oracle in the CPU test suite (no GPU in the dev container).
#include <stdint.h>
#include <string.h>
#define __restrict__
#include "../../metasbt_b200/csrc/msbt_kmer.h"

template <int KW>
static void run(const uint64_t *packed, const uint32_t *ends, uint32_t n_runs, uint32_t n_bases, uint32_t k,
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

extern "C" int emu_makebf(const uint64_t *packed, const uint32_t *ends, uint32_t n_runs, uint32_t n_bases, uint32_t k,
                          uint64_t seed, uint64_t modulus, uint64_t num_bits, uint64_t *words) {
    if (k < 1 || k > 64) return -1;
    memset(words, 0, ((num_bits + 63) / 64) * 8);
    msbt_pos_spec ps = msbt_make_pos_spec(modulus, num_bits, seed);
    if (k <= 32) run<1>(packed, ends, n_runs, n_bases, k, ps, (uint32_t *)words);
    else run<2>(packed, ends, n_runs, n_bases, k, ps, (uint32_t *)words);
    return 0;
}
