/*
 * msbt.h -- C-ABI of libmsbt.so, the B200 (sm_100a) replacement for the arithmetic that
 * MetaSBT obtains by spawning the external `howdesbt` binary.
 *
 * The reference has no FFI: its "operator interface" for this path is argv + stdout text of
 *     howdesbt makebf | bfoperate | bfdistance | dumpbf | query
 * launched with subprocess.check_call from /root/reference/metasbt/core.py.  Each entry
 * point below names the call site it replaces.  Signatures use plain pointers and sizes
 * only (no torch, no C++ types); device pointers are raw CUdeviceptr-compatible addresses
 * owned by the caller; `stream` is a cudaStream_t passed as void*.  Every function returns
 * 0 on success or a negative code (MSBT_E*), never throws, and is asynchronous on `stream`
 * unless stated otherwise.  msbt_last_error(ctx) gives the message for the last failure.
 *
 * Filter layout (device and `.bf` payload alike): ceil(num_bits/64) little-endian u64 words,
 * bit i = (word[i >> 6] >> (i & 63)) & 1 (sdsl::bit_vector).
 */
#ifndef MSBT_H
#define MSBT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MSBT_ABI_VERSION 2

#define MSBT_OK 0
#define MSBT_EINVAL (-22)  /* bad argument */
#define MSBT_ENOMEM (-12)  /* device or host allocation failed */
#define MSBT_ENOSPC (-28)  /* caller-provided buffer too small */
#define MSBT_ECUDA (-5)    /* CUDA runtime error; see msbt_last_error */

typedef struct msbt_ctx msbt_ctx;

/* Everything this path takes from HowDeSBT rather than from MetaSBT and that could not be pinned against a real
 * `howdesbt` binary (the binary and its sources are not part of the reference tree), as DATA: a context
 * created with another spec hashes, canonicalises and labels files that way without a rebuild -- the day a binary is at
 * hand, re-pinning is a configuration change (tests/pin_against_howdesbt.py finds the matching spec).  The default
 * (msbt_hash_spec_default) is App. A as written; it is the one the tuned K1 kernels have compiled in, any other spec is
 * served by a generic K1 kernel (min_abund >= 2 needs the default).  The oracle mirrors the same switches. */
#define MSBT_KEY_BYTES_CEIL_QUARTER 0u /* key = the low (k+3)/4 bytes of the packed canonical k-mer (default) */
#define MSBT_KEY_BYTES_WHOLE_WORDS 1u  /* key = whole 64-bit words: 8 bytes for k <= 32, 16 beyond */
#define MSBT_CANONICAL_MIN_FIRST_BASE_MSB 0u /* min(k-mer, reverse complement) as integers, first base most significant (default) */
#define MSBT_CANONICAL_FORWARD_ONLY 1u       /* the k-mer as read, no reverse complement */
typedef struct msbt_hash_spec {
    uint32_t struct_size;    /* sizeof(msbt_hash_spec), for forward compatibility */
    uint32_t key_bytes;      /* MSBT_KEY_BYTES_* : how many bytes of the packed k-mer MurmurHash3_x64_128 digests */
    uint32_t digest_half;    /* 0 = first 64 bits of the 128-bit digest (default), 1 = second */
    uint32_t canonical;      /* MSBT_CANONICAL_* */
    uint64_t bf_magic;       /* `.bf` header constants (App. A2); carried for the host-side file writer */
    uint32_t bf_version;
    uint32_t bf_header_size;
} msbt_hash_spec;
void msbt_hash_spec_default(msbt_hash_spec *spec);

/* One context per GPU per host thread.  Owns library scratch (tile tables, counters).  spec == NULL: the default spec.
 * A context must be driven from one stream at a time (its scratch and staging buffers are shared by its calls); every
 * entry point runs on the context's device and restores the caller's current device before it returns. */
int msbt_ctx_create(int device, const msbt_hash_spec *spec, msbt_ctx **out);
int msbt_ctx_get_spec(const msbt_ctx *ctx, msbt_hash_spec *out);
void msbt_ctx_destroy(msbt_ctx *ctx);
const char *msbt_last_error(const msbt_ctx *ctx);
int msbt_abi_version(void);
int msbt_sync(msbt_ctx *ctx, void *stream);
/* Number of kernels this context has launched so far (bench.py's gpu_launches). */
uint64_t msbt_launch_count(const msbt_ctx *ctx);

/* ---------------------------------------------------------------- host-side ingest (no GPU)
 * FASTA text -> 2-bit packed valid bases + run table.  Replaces howdesbt's own FASTA reader
 * (the file argument of `makebf`, core.py:3928, and of `query`, core.py:2046).
 * Records start at '>' lines, sequence lines are concatenated, whitespace is skipped, case is
 * folded, every other non-ACGT byte and every record boundary ends a run; runs shorter than
 * `min_run` (pass k) are dropped because they hold no k-mer.
 *   packed   : u64 words, 32 bases each, first base in the top two bits, zero padded, followed
 *              by MSBT_PACK_PAD_WORDS zero words; capacity in words
 *   run_ends : cumulative end offset (in bases of the packed stream) of each kept run
 * Call with packed == NULL to get the sizes only.  Synchronous, host only. */
#define MSBT_PACK_PAD_WORDS 3
int msbt_fasta_pack(const char *text, size_t n, uint32_t min_run,
                    uint64_t *packed, size_t packed_cap_words,
                    uint32_t *run_ends, size_t run_cap,
                    uint64_t *n_bases_out, uint64_t *n_runs_out);

/* The same packing ON THE DEVICE for a batch of FASTA files: `d_text` holds the files back to back,
 * h_offsets (host, n_files + 1 entries) delimits them.  Genome i is written at a position derived from the file sizes
 * only (every byte counted as a base, so nothing depends on the parse): packed words from
 * sum_{j<i} (ceil(len_j / 32) + MSBT_PACK_PAD_WORDS), run ends from sum_{j<i} (len_j / max(min_run, 1) + 1); the caller
 * sizes d_packed / d_run_ends accordingly (MSBT_ENOSPC otherwise).  h_descs (host, n_files entries) receives the
 * descriptors msbt_kmer_bloom_build takes (filter_index = i).  Result identical to msbt_fasta_pack per file.
 * Synchronises `stream` twice (the parse decides how many runs exist). */
struct msbt_genome_desc_s;
int msbt_fasta_pack_device(msbt_ctx *ctx, const char *d_text, const uint64_t *h_offsets, uint32_t n_files, uint32_t min_run,
                           uint64_t *d_packed, uint64_t packed_cap_words, uint32_t *d_run_ends, uint64_t run_cap,
                           struct msbt_genome_desc_s *h_descs, void *stream);

/* One genome of a construction / query batch. */
typedef struct msbt_genome_desc_s {
    uint64_t packed_word_off; /* first word of this genome inside d_packed */
    uint64_t n_bases;         /* valid bases (< 2^32) */
    uint64_t run_off;         /* first entry of this genome inside d_run_ends */
    uint32_t n_runs;
    uint32_t filter_index;    /* output filter = d_filters + filter_index * filter_stride_words */
} msbt_genome_desc;

/* ---------------------------------------------------------------- K1: howdesbt makebf
 * Replaces core.py:3920-3931 (Entry.sketch).  For every genome of the batch: every length-k
 * window inside a run -> canonical k-mer -> MurmurHash3_x64_128(seed) first half ->
 * pos = h % modulus -> set bit pos of its filter if pos < num_bits (--hashes=1).
 * min_abund <= 1: every k-mer; min_abund >= 2: only k-mers whose exact canonical multiplicity
 * inside the genome is >= min_abund (k <= 32 only in this version).
 * The output filters are fully written (zero-filled first) -- callers need not clear them.
 * `h_descs` is host memory and is consumed before the call returns.
 * variant: 0 = auto, 1 = direct (memset + red.or into HBM), 2 = tiled (partition kernel + tile-build kernel),
 *          3 = fused (one persistent warp-specialised kernel: partition warps + tile-build warps). */
int msbt_kmer_bloom_build(msbt_ctx *ctx, const uint64_t *d_packed, const uint32_t *d_run_ends,
                          const msbt_genome_desc *h_descs, uint32_t n_genomes,
                          uint32_t k, uint64_t seed, uint64_t modulus, uint64_t num_bits,
                          uint32_t min_abund, uint64_t *d_filters, uint64_t filter_stride_words,
                          int variant, void *stream);

/* ---------------------------------------------------------------- K2: howdesbt bfoperate --or
 * Replaces core.py:3851-3857 (Entry.index).  d_dst = OR of the n source filters.
 * `h_srcs` is a host array of n device pointers, consumed before the call returns. */
int msbt_bf_or_reduce(msbt_ctx *ctx, const uint64_t *const *h_srcs, uint32_t n, uint64_t words,
                      uint64_t *d_dst, void *stream);

/* ---------------------------------------------------------------- occurrence classes of a set of filters
 * One streaming pass over n filters: d_out[0] = bits set in at least one filter, d_out[1] = bits set in at least two.
 * Serves the k-mer size search that stands in for `kitsune kopt` (core.py:439-468): the number of distinct k-mers
 * shared by two or more genomes.  d_out: 2 u64 in device memory, zeroed by the call. */
int msbt_bf_occurrence(msbt_ctx *ctx, const uint64_t *const *h_filters, uint32_t n, uint64_t words,
                       uint64_t *d_out, void *stream);

/* ---------------------------------------------------------------- K3': howdesbt dumpbf --show:density
 * Replaces core.py:3588-3593 (Entry.get_density).  d_out[i] = popcount(filter i). */
int msbt_bf_popcount(msbt_ctx *ctx, const uint64_t *const *h_filters, uint32_t n, uint64_t words,
                     uint64_t *d_out, void *stream);

/* ---------------------------------------------------------------- K3: howdesbt bfdistance
 * Replaces BOTH spawns of core.py:1747-1753 (--show:intersect) and :1768-1774 (--show:union)
 * (Database.dist) with one pass: d_intersect[j] = popc(focus & target_j),
 * d_union[j] = popc(focus | target_j). */
int msbt_bf_and_popcount_1vN(msbt_ctx *ctx, const uint64_t *d_focus,
                             const uint64_t *const *h_targets, uint32_t n, uint64_t words,
                             uint64_t *d_intersect, uint64_t *d_union, void *stream);

/* An arbitrary list of pairs in one launch: d_intersect[p] = popc(a_p & b_p), d_union[p] = popc(a_p | b_p).
 * The batch form of Database.dist for `profile_many`: every MAG of a batch against its own few cluster
 * representatives / species centroids (core.py:2106-2198 calls dist once per MAG and level).
 * h_a / h_b are host arrays of n_pairs device pointers, consumed before the call returns. */
int msbt_bf_and_popcount_pairs(msbt_ctx *ctx, const uint64_t *const *h_a, const uint64_t *const *h_b, uint32_t n_pairs,
                               uint64_t words, uint64_t *d_intersect, uint64_t *d_union, void *stream);

/* All pairs among n filters (Entry.get_boundaries, core.py:4036-4081, which calls dist once
 * per source).  d_intersect is a dense n*n row-major u64 matrix; the diagonal holds
 * popcount(filter i), so union(i,j) = I[i][i] + I[j][j] - I[i][j]. */
int msbt_bf_and_popcount_allpairs(msbt_ctx *ctx, const uint64_t *const *h_filters, uint32_t n,
                                  uint64_t words, uint64_t *d_intersect, void *stream);

/* Every filter of list A against every filter of list B: d_intersect is a dense n_a x n_b row-major u64 matrix of
 * popc(a_i & b_j).  The building block of an all-pairs matrix that does not fit one GPU: every rank holds a block of the
 * filters and meets the other blocks one after the other (shard.all_pairs_ring). */
int msbt_bf_and_popcount_rect(msbt_ctx *ctx, const uint64_t *const *h_a, uint32_t n_a,
                              const uint64_t *const *h_b, uint32_t n_b, uint64_t words,
                              uint64_t *d_intersect, void *stream);

/* ---------------------------------------------------------------- K4: howdesbt query --distinctkmers
 * Replaces core.py:2039-2047 (Database.profile).
 * Step 1: the query's distinct positions are exactly the set bits of its own min=1 filter
 * (App. A6 corollary), so they are produced by K1 + this ordered extraction:
 * d_positions (u32, ascending) receives at most `cap` entries, *d_count (u64) the true count. */
int msbt_bf_extract_positions(msbt_ctx *ctx, const uint64_t *d_filter, uint64_t words,
                              uint32_t *d_positions, uint64_t cap, uint64_t *d_count, void *stream);

/* The same for a batch of n query filters in one call and without a host round trip: d_offsets (u64, n + 1 entries)
 * receives the start of every query inside d_positions (offsets[n] = total count); positions beyond `cap` are dropped
 * (call once with cap = 0 to size the buffer, or pass an upper bound such as the summed k-mer counts). */
int msbt_bf_extract_positions_batch(msbt_ctx *ctx, const uint64_t *const *h_filters, uint32_t n, uint64_t words,
                                    uint32_t *d_positions, uint64_t cap, uint64_t *d_offsets, void *stream);

/* HOST side of a sparse filter transfer (no GPU, synchronous): the ascending positions msbt_bf_extract_positions*
 * produced, copied device -> host by the caller, -> the dense filter `words` (ceil(num_bits/64) u64, fully written).
 * A 0.5 %-dense filter travels as 4 bytes per set bit instead of num_bits/8 bytes (makebf's output side, core.py:3929
 * `--out=<bf>`).  MSBT_EINVAL if the positions are not ascending or reach num_bits. */
int msbt_positions_expand(const uint32_t *positions, uint64_t n, uint64_t num_bits, uint64_t *words);

/* Step 2: hits[q][l] = #{p in positions of query q : bit p of leaf filter l is set}.
 * d_positions holds all queries back to back; h_query_offsets (host, n_queries+1 entries)
 * delimits them.  d_hits is n_queries * n_leaves u32, row-major. */
int msbt_query_batch(msbt_ctx *ctx, const uint64_t *const *h_leaves, uint32_t n_leaves,
                     const uint32_t *d_positions, const uint64_t *h_query_offsets,
                     uint32_t n_queries, uint32_t *d_hits, void *stream);

/* The same with an explicit [begin, end) range of d_positions per query (h_query_ranges: 2 * n_queries entries,
 * begin_0, end_0, begin_1, ...): the queries of a call may be any subset, in any order, of a packed batch of position
 * lists -- `profile_many` probes every flat tree with just the MAGs that reached it, without copying their positions. */
int msbt_query_batch_ranges(msbt_ctx *ctx, const uint64_t *const *h_leaves, uint32_t n_leaves,
                            const uint32_t *d_positions, const uint64_t *h_query_ranges,
                            uint32_t n_queries, uint32_t *d_hits, void *stream);

/* Bit-sliced leaf matrix for large leaf counts: row p holds bit p of every
 * leaf, msbt_sliced_row_words(n_leaves) u32 words per row (ceil(n_leaves/32) rounded up to a
 * multiple of 8 words, so that rows start on 32-byte sectors; the content of the padding words is ignored).
 * Rows [row_begin, row_end) only, so that a rank that owns a bit range builds and probes just
 * its shard: d_sliced holds (row_end - row_begin) * msbt_sliced_row_words(n_leaves) words. */
uint32_t msbt_sliced_row_words(uint32_t n_leaves);
int msbt_bitslice_build(msbt_ctx *ctx, const uint64_t *const *h_leaves, uint32_t n_leaves,
                        uint64_t row_begin, uint64_t row_end, uint32_t *d_sliced, void *stream);

/* Same result as msbt_query_batch restricted to positions in [row_begin, row_end); partial
 * hit matrices of different ranges add up to the full one (NCCL sum across ranks). */
int msbt_query_batch_sliced(msbt_ctx *ctx, const uint32_t *d_sliced, uint32_t n_leaves,
                            uint64_t row_begin, uint64_t row_end,
                            const uint32_t *d_positions, const uint64_t *h_query_offsets,
                            uint32_t n_queries, uint32_t *d_hits, void *stream);

/* msbt_query_batch_sliced with explicit per-query ranges (see msbt_query_batch_ranges). */
int msbt_query_batch_sliced_ranges(msbt_ctx *ctx, const uint32_t *d_sliced, uint32_t n_leaves,
                                   uint64_t row_begin, uint64_t row_end,
                                   const uint32_t *d_positions, const uint64_t *h_query_ranges,
                                   uint32_t n_queries, uint32_t *d_hits, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* MSBT_H */
