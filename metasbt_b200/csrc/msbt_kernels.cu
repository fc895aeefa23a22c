// msbt_kernels.cu -- sm_100a kernels + C-ABI (include/msbt.h) for the MetaSBT hot path.
//
// What each kernel replaces (all in /root/reference/metasbt/core.py):
//   kmer_bloom_*      howdesbt makebf       :3920-3931
//   bf_or_reduce      howdesbt bfoperate    :3851-3857
//   bf_popcount       howdesbt dumpbf       :3588-3593
//   bf_and_popcount_* howdesbt bfdistance   :1747-1753, :1768-1774 (and the loop :4036-4081)
//   bf_extract / query_* howdesbt query     :2039-2047
//
// None of this is a dense contraction: every kernel is HBM/L2-bound integer work, so there
// are no tensor-core instructions here by design (BASELINE.json north_star).  The rules that
// matter are coalesced 128-bit access, one scattered memory operation per k-mer at most, and
// persistent grids sized in multiples of the SM count.
#include <cuda_runtime.h>
#include <errno.h>
#include <stdarg.h>
#include <stdio.h>
#include <math.h>
#include <string.h>
#include <emmintrin.h>

#include <algorithm>
#include <vector>

#include "../../include/msbt.h"
#include "msbt_spec.h"
#include "msbt_kmer.h"

// --------------------------------------------------------------------------- context

struct msbt_ctx {
    int device;
    int num_sms;
    char err[512];
    uint64_t launches;
    // library-owned device scratch, grown on demand
    void *d_scratch;
    size_t scratch_bytes;
    // pinned staging for small host tables handed to kernels
    void *h_stage;
    size_t stage_bytes;
    cudaEvent_t stage_free;  // recorded after the last H2D that read h_stage
    bool stage_busy;
    msbt_hash_spec spec;
    bool spec_is_default;
    bool tiled_attr_set;  // dynamic-smem opt-in done for this device
    bool fused_attr_set;
    bool phased_attr_set;
    // side streams of the `--min >= 2` build: consecutive genomes run concurrently (their passes are bound by different
    // units: L2 atomics, issue slots, latency), forked from and joined to the caller's stream inside the call
    cudaStream_t side[4];
    cudaEvent_t side_done[4];
    cudaEvent_t fork_ev;
    int n_side;
};

static int fail(msbt_ctx *ctx, int code, const char *fmt, ...) {
    if (ctx) {
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(ctx->err, sizeof(ctx->err), fmt, ap);
        va_end(ap);
    }
    return code;
}

#define CU_TRY(ctx, call)                                                                  \
    do {                                                                                   \
        cudaError_t e__ = (call);                                                          \
        if (e__ != cudaSuccess)                                                            \
            return fail(ctx, MSBT_ECUDA, "%s failed: %s", #call, cudaGetErrorString(e__)); \
    } while (0)

// Every entry point runs on the context's device and puts the caller's current device back on return (a process that
// drives several contexts, or whose torch current device is another GPU, keeps its own bookkeeping intact).
struct DeviceGuard {
    int prev = -1;
    bool changed = false;
    cudaError_t enter(int device) {
        cudaError_t e = cudaGetDevice(&prev);
        if (e != cudaSuccess) return e;
        if (prev == device) return cudaSuccess;
        e = cudaSetDevice(device);
        changed = e == cudaSuccess;
        return e;
    }
    ~DeviceGuard() {
        if (changed) cudaSetDevice(prev);
    }
};
#define CTX_DEVICE(ctx)      \
    DeviceGuard dev_guard__; \
    CU_TRY(ctx, dev_guard__.enter((ctx)->device))

static int ensure_scratch(msbt_ctx *ctx, size_t bytes) {
    if (bytes <= ctx->scratch_bytes) return MSBT_OK;
    if (ctx->d_scratch) {
        CU_TRY(ctx, cudaDeviceSynchronize());
        CU_TRY(ctx, cudaFree(ctx->d_scratch));
        ctx->d_scratch = nullptr;
        ctx->scratch_bytes = 0;
    }
    size_t want = bytes + bytes / 4 + 4096;
    if (cudaMalloc(&ctx->d_scratch, want) != cudaSuccess) {
        cudaGetLastError();
        return fail(ctx, MSBT_ENOMEM, "cudaMalloc(%zu) for scratch failed", want);
    }
    ctx->scratch_bytes = want;
    return MSBT_OK;
}

// Copy a small host table to device scratch at `d_dst` through pinned staging, asynchronously
// on `stream`.  The staging buffer is recycled only after the previous copy has finished.
static int stage_to_device(msbt_ctx *ctx, void *d_dst, const void *h_src, size_t bytes, size_t stage_off,
                           cudaStream_t stream) {
    memcpy((char *)ctx->h_stage + stage_off, h_src, bytes);
    CU_TRY(ctx, cudaMemcpyAsync(d_dst, (char *)ctx->h_stage + stage_off, bytes, cudaMemcpyHostToDevice, stream));
    return MSBT_OK;
}

static int stage_begin(msbt_ctx *ctx, size_t bytes) {
    if (ctx->stage_busy) {
        CU_TRY(ctx, cudaEventSynchronize(ctx->stage_free));
        ctx->stage_busy = false;
    }
    if (bytes > ctx->stage_bytes) {
        if (ctx->h_stage) CU_TRY(ctx, cudaFreeHost(ctx->h_stage));
        ctx->h_stage = nullptr;
        ctx->stage_bytes = 0;
        size_t want = bytes + bytes / 2 + 4096;
        if (cudaMallocHost(&ctx->h_stage, want) != cudaSuccess) {
            cudaGetLastError();
            return fail(ctx, MSBT_ENOMEM, "cudaMallocHost(%zu) failed", want);
        }
        ctx->stage_bytes = want;
    }
    return MSBT_OK;
}

static int stage_end(msbt_ctx *ctx, cudaStream_t stream) {
    CU_TRY(ctx, cudaEventRecord(ctx->stage_free, stream));
    ctx->stage_busy = true;
    return MSBT_OK;
}

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

extern "C" int msbt_abi_version(void) { return MSBT_ABI_VERSION; }

extern "C" void msbt_hash_spec_default(msbt_hash_spec *spec) {
    if (!spec) return;
    memset(spec, 0, sizeof(*spec));
    spec->struct_size = (uint32_t)sizeof(*spec);
    spec->key_bytes = MSBT_KEY_BYTES_CEIL_QUARTER;
    spec->digest_half = 0;
    spec->canonical = MSBT_CANONICAL_MIN_FIRST_BASE_MSB;
    spec->bf_magic = MSBT_BF_MAGIC;
    spec->bf_version = MSBT_BF_VERSION;
    spec->bf_header_size = MSBT_BF_HEADER_SIZE;
}

extern "C" int msbt_ctx_get_spec(const msbt_ctx *ctx, msbt_hash_spec *out) {
    if (!ctx || !out) return MSBT_EINVAL;
    *out = ctx->spec;
    return MSBT_OK;
}

extern "C" int msbt_ctx_create(int device, const msbt_hash_spec *spec, msbt_ctx **out) {
    if (!out) return MSBT_EINVAL;
    *out = nullptr;
    if (spec && (spec->struct_size != sizeof(msbt_hash_spec) || spec->key_bytes > 1 || spec->digest_half > 1 || spec->canonical > 1)) {
        fprintf(stderr, "msbt_ctx_create: invalid msbt_hash_spec\n");
        return MSBT_EINVAL;
    }
    msbt_ctx *ctx = new (std::nothrow) msbt_ctx();
    if (!ctx) return MSBT_ENOMEM;
    memset(ctx, 0, sizeof(*ctx));
    ctx->device = device;
    msbt_hash_spec_default(&ctx->spec);
    if (spec) ctx->spec = *spec;
    ctx->spec_is_default = ctx->spec.key_bytes == MSBT_KEY_BYTES_CEIL_QUARTER && ctx->spec.digest_half == 0 &&
                           ctx->spec.canonical == MSBT_CANONICAL_MIN_FIRST_BASE_MSB;
    DeviceGuard guard;
    cudaError_t e = guard.enter(device);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&ctx->num_sms, cudaDevAttrMultiProcessorCount, device);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->stage_free, cudaEventDisableTiming);
    if (e != cudaSuccess) {
        // No CPU fallback exists: without a usable GPU the library refuses to construct.
        fprintf(stderr, "msbt_ctx_create(device=%d): %s\n", device, cudaGetErrorString(e));
        delete ctx;
        return MSBT_ECUDA;
    }
    *out = ctx;
    return MSBT_OK;
}

extern "C" void msbt_ctx_destroy(msbt_ctx *ctx) {
    if (!ctx) return;
    DeviceGuard guard;
    guard.enter(ctx->device);
    cudaDeviceSynchronize();
    if (ctx->d_scratch) cudaFree(ctx->d_scratch);
    if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
    if (ctx->stage_free) cudaEventDestroy(ctx->stage_free);
    for (int i = 0; i < ctx->n_side; i++) {
        cudaStreamDestroy(ctx->side[i]);
        cudaEventDestroy(ctx->side_done[i]);
    }
    if (ctx->fork_ev) cudaEventDestroy(ctx->fork_ev);
    delete ctx;
}

extern "C" const char *msbt_last_error(const msbt_ctx *ctx) { return ctx ? ctx->err : "null context"; }
extern "C" uint64_t msbt_launch_count(const msbt_ctx *ctx) { return ctx ? ctx->launches : 0; }

extern "C" int msbt_sync(msbt_ctx *ctx, void *stream) {
    if (!ctx) return MSBT_EINVAL;
    CU_TRY(ctx, cudaStreamSynchronize((cudaStream_t)stream));
    return MSBT_OK;
}

#define LAUNCH_CHECK(ctx)                 \
    do {                                  \
        (ctx)->launches++;                \
        CU_TRY(ctx, cudaGetLastError());  \
    } while (0)

// --------------------------------------------------------------------------- host FASTA ingest

extern "C" int msbt_fasta_pack(const char *text, size_t n, uint32_t min_run, uint64_t *packed,
                               size_t packed_cap_words, uint32_t *run_ends, size_t run_cap,
                               uint64_t *n_bases_out, uint64_t *n_runs_out) {
    if (!text && n) return MSBT_EINVAL;
    if (min_run > 64) return MSBT_EINVAL;
    static const struct Lut {
        uint8_t v[256];
        Lut() {
            for (int i = 0; i < 256; i++) v[i] = MSBT_BASE_INVALID;
            v['A'] = v['a'] = 0; v['C'] = v['c'] = 1; v['G'] = v['g'] = 2; v['T'] = v['t'] = 3;
            v[' '] = v['\t'] = v['\r'] = 5;  // skipped, does not break a run
            v['\n'] = 6;
        }
    } lut;
    const bool write = packed != nullptr;
    uint64_t nb = 0;       // bases committed to the packed stream
    uint64_t nruns = 0;
    uint64_t cur = 0;      // word being filled
    uint32_t run_len = 0;  // length of the current run, saturating at min_run once committed
    uint8_t pending[64];   // a run is committed only once it has reached min_run bases
    int rc = MSBT_OK;
    auto commit = [&](uint8_t v) -> bool {
        if (write) {
            const uint64_t w = nb >> 5;
            if (w + MSBT_PACK_PAD_WORDS >= packed_cap_words) return false;
            cur |= (uint64_t)v << (62 - 2 * (nb & 31));
            if ((nb & 31) == 31) { packed[w] = cur; cur = 0; }
        }
        nb++;
        return true;
    };
    auto end_run = [&]() {
        if (run_len >= min_run && run_len > 0) {
            if (write) {
                if (nruns < run_cap) run_ends[nruns] = (uint32_t)nb; else rc = MSBT_ENOSPC;
            }
            nruns++;
        }
        run_len = 0;
    };
    bool line_start = true;
    for (size_t i = 0; i < n;) {
        const unsigned char c = (unsigned char)text[i];
        if (line_start && c == '>') {
            const void *nl = memchr(text + i, '\n', n - i);
            i = nl ? (size_t)((const char *)nl - text) : n;
            end_run();
            continue;
        }
        i++;
        const uint8_t v = lut.v[c];
        if (v < 4) {
            line_start = false;
            if (run_len + 1 < min_run) {
                pending[run_len++] = v;
            } else {
                if (run_len + 1 == min_run) {  // the run just became long enough: flush what was held back
                    for (uint32_t t = 0; t < run_len; t++)
                        if (!commit(pending[t])) return MSBT_ENOSPC;
                }
                if (!commit(v)) return MSBT_ENOSPC;
                if (run_len < 0xFFFFFFFFu) run_len++;
                if (nb >> 32) return MSBT_EINVAL;  // run table is u32
            }
        } else if (v == 6) {
            line_start = true;
        } else if (v == 5) {
            line_start = false;
        } else {
            line_start = false;
            end_run();
        }
    }
    end_run();
    if (write) {
        const uint64_t w = (nb + 31) >> 5;
        if (w + MSBT_PACK_PAD_WORDS > packed_cap_words) return MSBT_ENOSPC;
        if (nb & 31) packed[nb >> 5] = cur;
        for (uint64_t p = 0; p < MSBT_PACK_PAD_WORDS; p++) packed[w + p] = 0;
    }
    if (n_bases_out) *n_bases_out = nb;
    if (n_runs_out) *n_runs_out = nruns;
    return rc;
}

// --------------------------------------------------------------------------- K1: k-mer -> Bloom

// One tile = one CTA iteration = TILE_THREADS threads x 32 k-mer starts (one packed word each).
constexpr int TILE_THREADS = 256;
constexpr uint32_t TILE_BASES = TILE_THREADS * 32;

struct DevGenome {
    uint64_t packed_word_off;
    uint64_t run_off;
    uint64_t filter_word_off;  // in u64 words
    uint32_t n_bases;
    uint32_t n_runs;
    uint32_t tile_begin;    // first global tile index of this genome (TILE_BASES units)
    uint32_t fchunk_begin;  // first chunk of this genome in the fused kernel (FU_CHUNK_BASES units)
};

struct KmerSpec {
    msbt_pos_spec pos;
    uint32_t k;
};

__device__ __forceinline__ uint32_t find_genome(const DevGenome *__restrict__ g, uint32_t n, uint32_t tile) {
    uint32_t lo = 0, hi = n;  // last genome with tile_begin <= tile
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (g[mid].tile_begin <= tile) lo = mid; else hi = mid;
    }
    return lo;
}

// Variant 1 ("direct"): filters are zero-filled by a preceding memset; every k-mer issues one
// RED.OR.b32 straight to its filter word.  One scattered memory operation per k-mer.
#define MSBT_DIRECT_STEP(J)                                                                                  \
    {                                                                                                        \
        uint64_t lo_, hi_;                                                                                   \
        kb.template get<J>(lo_, hi_);                                                                        \
        const uint64_t pos_ = msbt_position_v<POW2>(msbt_kmer_hash_v<HV>(lo_, hi_, len, spec.pos.seed), spec.pos); \
        if (((vg >> J) & 1u) && pos_ < spec.pos.num_bits) atomicOr(out + (pos_ >> 5), 1u << (pos_ & 31));    \
    }

template <int KW, int HV, bool POW2>
__global__ void __launch_bounds__(TILE_THREADS, 4)
kmer_bloom_direct_kernel(const uint64_t *__restrict__ packed, const uint32_t *__restrict__ run_ends,
                         const DevGenome *__restrict__ genomes, uint32_t n_genomes, uint32_t n_tiles,
                         KmerSpec spec, uint32_t *__restrict__ filters, const uint32_t *__restrict__ only_dirty,
                         uint32_t tile_base = 0) {
    const uint32_t k = spec.k, len = (k + 3) >> 2;
    if (only_dirty) {  // redo pass: leave at once when no genome is marked (the normal case)
        int any = 0;
        for (uint32_t i = threadIdx.x; i < n_genomes; i += blockDim.x) any |= (only_dirty[i] != 0);
        if (!__syncthreads_or(any)) return;
    }
    // tile_base / n_tiles: the (absolute) tile range of this launch -- a group of genomes whose filters fit L2 together
    for (uint32_t tile = tile_base + blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const uint32_t gi = find_genome(genomes, n_genomes, tile);
        const DevGenome g = genomes[gi];
        // redo pass after the fused kernel: only genomes whose ring buckets overflowed
        if (only_dirty) {
            if (only_dirty[gi] == 0) {
                const uint32_t nxt = (gi + 1 < n_genomes) ? genomes[gi + 1].tile_begin : n_tiles;  // skip to the next genome
                const uint32_t skip = (nxt - tile + gridDim.x - 1) / gridDim.x;
                tile += (skip ? skip - 1 : 0) * gridDim.x;
                continue;
            }
        }
        const uint32_t w = (tile - g.tile_begin) * TILE_THREADS + threadIdx.x;
        const uint32_t s0 = w * 32;
        if (s0 >= g.n_bases) continue;
        const uint32_t vm = valid_mask(run_ends + g.run_off, g.n_runs, s0, k);
        if (vm == 0) continue;
        KmerBlock<KW> kb;
        kb.init(packed + g.packed_word_off + w, k);
        uint32_t *__restrict__ out = filters + 2 * g.filter_word_off;
#pragma unroll 1
        for (uint32_t grp = 0; grp < 4; grp++) {
            const uint32_t vg = (vm >> (8 * grp)) & 0xFFu;
            MSBT_DIRECT_STEP(0) MSBT_DIRECT_STEP(1) MSBT_DIRECT_STEP(2) MSBT_DIRECT_STEP(3)
            MSBT_DIRECT_STEP(4) MSBT_DIRECT_STEP(5) MSBT_DIRECT_STEP(6) MSBT_DIRECT_STEP(7)
            kb.advance8();
        }
    }
}

// Generic K1 kernel for contexts with a non-default msbt_hash_spec: the direct kernel's structure (memset + one RED per
// k-mer) with the key length, the digest half and the canonical rule as run-time arguments.  Correctness path for
// re-pinning against a real howdesbt; the tuned kernels keep the default spec compiled in.
template <int KW>
__global__ void __launch_bounds__(TILE_THREADS, 4)
kmer_bloom_generic_kernel(const uint64_t *__restrict__ packed, const uint32_t *__restrict__ run_ends,
                          const DevGenome *__restrict__ genomes, uint32_t n_genomes, uint32_t n_tiles, KmerSpec spec,
                          uint32_t key_len, uint32_t second_half, uint32_t forward_only, uint32_t *__restrict__ filters) {
    const uint32_t k = spec.k;
    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const DevGenome g = genomes[find_genome(genomes, n_genomes, tile)];
        const uint32_t w = (tile - g.tile_begin) * TILE_THREADS + threadIdx.x;
        const uint32_t s0 = w * 32;
        if (s0 >= g.n_bases) continue;
        const uint32_t vm = valid_mask(run_ends + g.run_off, g.n_runs, s0, k);
        if (vm == 0) continue;
        KmerWalker<KW> kw;
        kw.init(packed + g.packed_word_off + w, k);
        uint32_t *__restrict__ out = filters + 2 * g.filter_word_off;
        for (uint32_t j = 0; j < 32; j++) {
            if ((vm >> j) & 1u) {
                uint64_t lo, hi;
                if (forward_only) { lo = kw.fl; hi = kw.fh; }
                else kw.canonical(lo, hi);
                const uint64_t pos = msbt_position(msbt_kmer_hash_generic(lo, hi, key_len, spec.pos.seed, second_half), spec.pos);
                if (pos < spec.pos.num_bits) atomicOr(out + (pos >> 5), 1u << (pos & 31));
            }
            kw.roll(j);
        }
    }
}

// ---- Variant 2 ("tiled"): no memset, no read-modify-write in HBM.
//   phase 1  kmer_partition_kernel : hash every k-mer, bucket its position by filter range.
//            Per CTA chunk the buckets are staged in shared memory (one smem atomicAdd per
//            k-mer), then flushed with one global reservation per (chunk, bucket) and
//            sector-sized coalesced stores into an L2-resident scratch.
//   phase 2  tile_build_kernel     : one CTA per 64 KiB filter tile: zero the tile in shared
//            memory, atomicOr its bucket's positions into it (smem atomics run at ~950 Gop/s
//            chip-wide, 5x the L2 RED rate -- profiles/r01_microbench_scatter.txt), then write
//            the finished tile to HBM once with a TMA bulk store (cp.async.bulk).
//   phase 3  overflow_fixup_kernel : positions that did not fit a staging slot (skewed inputs)
//            are OR-ed in afterwards with RED; normally a handful per genome.
// HBM traffic per genome = packed input + filter written once (+ the scratch if it spills L2).
constexpr int PT_THREADS = 256;
constexpr int PT_CAP = 16;                  // staged entries per bucket per chunk (expected <= 8)
constexpr uint32_t PT_MAX_BUCKETS = 1024;   // buckets per filter
constexpr uint32_t PT_CNT_WORDS = PT_MAX_BUCKETS;
constexpr size_t PT_SMEM_BYTES = ((size_t)2 * PT_CNT_WORDS + (size_t)(PT_MAX_BUCKETS + 1) * 16) * 4;
constexpr uint32_t TB_TILE_LOG2 = 19;       // 2^19 bits = 64 KiB per tile, 3 CTAs per SM
constexpr uint32_t TB_TILE_WORDS32 = 1u << (TB_TILE_LOG2 - 5);
constexpr int TB_THREADS = 256;

struct TiledParams {
    uint32_t g0;                  // first genome of the group in the DevGenome table
    uint32_t n_group;             // genomes in the group
    uint32_t tile0;               // input-tile index (whole batch) of the group's first chunk
    uint32_t n_chunks;            // input chunks in the group
    uint32_t bucket_shift;        // bucket = pos >> bucket_shift  (>= TB_TILE_LOG2)
    uint32_t buckets_per_filter;  // <= PT_MAX_BUCKETS
    uint32_t tiles_per_filter;
    uint32_t gcap;                // scratch entries per bucket
    uint64_t words;               // u64 words per filter
    uint64_t ovf_cap;
};

__device__ __forceinline__ void overflow_push(unsigned long long *ovf_count, uint64_t *ovf, uint64_t cap, uint64_t bit) {
    unsigned long long i = atomicAdd(ovf_count, 1ull);
    if (i < cap) ovf[i] = bit;
}

__device__ __noinline__ void overflow_push_cold(unsigned long long *ovf_count, uint64_t *ovf, uint64_t cap, uint64_t bit) {
    overflow_push(ovf_count, ovf, cap, bit);
}

// One k-mer of the unrolled group, split in three sweeps so that the eight hash chains of a group
// are independent straight-line ALU work (ILP), the eight shared-memory atomics are issued back to
// back (their latencies overlap), and only then the eight dependent stores.  Positions are 32-bit
// here (the tiled path requires num_bits <= 2^32).  K-mers that must not be staged (outside a run,
// or pos >= num_bits) are counted into a dummy bucket NB so that the hot path has no branch.
#define MSBT_PT_HASH(J)                                                                                      \
    {                                                                                                        \
        uint64_t lo_, hi_;                                                                                   \
        kb.template get<J>(lo_, hi_);                                                                        \
        const uint64_t p_ = msbt_position_v<POW2>(msbt_kmer_hash_v<HV>(lo_, hi_, len, spec.pos.seed), spec.pos); \
        pos8[J] = (uint32_t)p_;                                                                              \
        const bool ok_ = ((vg >> J) & 1u) && (FULL || p_ < spec.pos.num_bits);                               \
        ok8[J] = ok_;                                                                                        \
        b8[J] = ok_ ? (pos8[J] >> P.bucket_shift) : NB;                                                      \
    }
#define PT_SKIP 0xFFFFFFFFu
// Predicated shared-memory atomic / store in PTX so that skipped k-mers cost a predicate, not a branch.
__device__ __forceinline__ uint32_t atoms_inc_if(uint32_t saddr, bool ok) {
    uint32_t r;
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\tmov.u32 %0, 0;\n\t@p atom.shared.add.u32 %0, [%1], 1;\n\t}"
                 : "=r"(r) : "r"(saddr), "r"((uint32_t)ok) : "memory");
    return r;
}
__device__ __forceinline__ void sts_if(uint32_t saddr, uint32_t v, bool ok) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p st.shared.u32 [%0], %1;\n\t}" ::"r"(saddr), "r"(v), "r"((uint32_t)ok) : "memory");
}
#define MSBT_PT_RANK(J) r8[J] = atoms_inc_if(cnt_s + b8[J] * 4u, ok8[J]);
#define MSBT_PT_STORE(J)                                                          \
    sts_if(stage_s + (b8[J] * PT_CAP + r8[J]) * 4u, pos8[J], r8[J] < PT_CAP);     \
    rmax = max(rmax, r8[J]);
#define MSBT_PT_OVERFLOW(J) \
    if (r8[J] >= PT_CAP) overflow_push_cold(ovf_count, ovf, P.ovf_cap, g.filter_word_off * 64 + pos8[J]);

// FULL: modulus <= num_bits, i.e. every position is in range and the `pos < num_bits` test is dropped.
template <int KW, int HV, bool POW2, bool FULL>
__global__ void __launch_bounds__(PT_THREADS, 3)
kmer_partition_kernel(const uint64_t *__restrict__ packed, const uint32_t *__restrict__ run_ends,
                      const DevGenome *__restrict__ genomes, TiledParams P, KmerSpec spec,
                      uint32_t *__restrict__ gcnt, uint32_t *__restrict__ gbuf,
                      unsigned long long *__restrict__ ovf_count, uint64_t *__restrict__ ovf) {
    extern __shared__ __align__(16) uint32_t pt_smem[];
    const uint32_t NB = P.buckets_per_filter;
    // fixed layout (compile-time offsets keep the staging address arithmetic to one IMAD):
    uint32_t *cnt = pt_smem;                    // [NB] entries staged per bucket
    uint32_t *base = pt_smem + PT_CNT_WORDS;    // [NB] reserved offset in the global bucket
    uint32_t *stage = pt_smem + 2 * PT_CNT_WORDS;  // [NB + 1][PT_CAP]; row NB is a dummy for skipped k-mers
    const uint32_t cnt_s = (uint32_t)__cvta_generic_to_shared(cnt);
    const uint32_t stage_s = (uint32_t)__cvta_generic_to_shared(stage);
    const uint32_t k = spec.k, len = (k + 3) >> 2;
    for (uint32_t chunk = blockIdx.x; chunk < P.n_chunks; chunk += gridDim.x) {
        const uint32_t tile = P.tile0 + chunk;
        const uint32_t gl = find_genome(genomes + P.g0, P.n_group, tile);
        const DevGenome g = genomes[P.g0 + gl];
        for (uint32_t b = threadIdx.x; b < NB; b += PT_THREADS) cnt[b] = 0;
        __syncthreads();
        const uint32_t w = (tile - g.tile_begin) * PT_THREADS + threadIdx.x;
        const uint32_t s0 = w * 32;
        uint32_t vm = 0;
        if (s0 < g.n_bases) vm = valid_mask(run_ends + g.run_off, g.n_runs, s0, k);
        if (vm) {
            KmerBlock<KW> kb;
            kb.init(packed + g.packed_word_off + w, k);
#pragma unroll 1
            for (uint32_t grp = 0; grp < 4; grp++) {
                const uint32_t vg = (vm >> (8 * grp)) & 0xFFu;
                uint32_t pos8[8], b8[8], r8[8];
                bool ok8[8];
                MSBT_PT_HASH(0) MSBT_PT_HASH(1) MSBT_PT_HASH(2) MSBT_PT_HASH(3)
                MSBT_PT_HASH(4) MSBT_PT_HASH(5) MSBT_PT_HASH(6) MSBT_PT_HASH(7)
                MSBT_PT_RANK(0) MSBT_PT_RANK(1) MSBT_PT_RANK(2) MSBT_PT_RANK(3)
                MSBT_PT_RANK(4) MSBT_PT_RANK(5) MSBT_PT_RANK(6) MSBT_PT_RANK(7)
                uint32_t rmax = 0;
                MSBT_PT_STORE(0) MSBT_PT_STORE(1) MSBT_PT_STORE(2) MSBT_PT_STORE(3)
                MSBT_PT_STORE(4) MSBT_PT_STORE(5) MSBT_PT_STORE(6) MSBT_PT_STORE(7)
                if (rmax >= PT_CAP) {  // rare: a staging row is full
                    MSBT_PT_OVERFLOW(0) MSBT_PT_OVERFLOW(1) MSBT_PT_OVERFLOW(2) MSBT_PT_OVERFLOW(3)
                    MSBT_PT_OVERFLOW(4) MSBT_PT_OVERFLOW(5) MSBT_PT_OVERFLOW(6) MSBT_PT_OVERFLOW(7)
                }
                kb.advance8();
            }
        }
        __syncthreads();
        // Flush: every (chunk, bucket) reserves a multiple of 4 entries so that the copy is one aligned
        // 16-byte load + store per quad; the pad slots repeat the bucket's first entry (OR is idempotent).
        // base[b] = PT_SKIP marks a bucket whose scratch is full: its entries take the overflow path.
        uint32_t *gc = gcnt + (uint64_t)gl * NB;
        for (uint32_t b = threadIdx.x; b < NB; b += PT_THREADS) {
            const uint32_t n = min(cnt[b], (uint32_t)PT_CAP);
            cnt[b] = n;
            uint32_t o = 0;
            if (n) {
                const uint32_t n4 = (n + 3u) & ~3u;
                const uint32_t o_raw = atomicAdd(gc + b, n4);
                if (o_raw + n4 <= P.gcap) {
                    o = b * P.gcap + o_raw;
                } else {
                    // scratch bucket full: the entries take the overflow path.  tile_build reads min(count, gcap) entries, so
                    // the tail this reservation leaves unwritten is padded with a position of this very bucket (OR is
                    // idempotent); stale scratch content there could set foreign bits.
                    o = PT_SKIP;
                    const uint32_t first = stage[b * PT_CAP];
                    uint32_t *tail = gbuf + ((uint64_t)gl * NB + b) * P.gcap;
                    for (uint32_t t = o_raw; t < P.gcap; t++) tail[t] = first;
                }
            }
            base[b] = o;
        }
        __syncthreads();
        uint4 *gb4 = (uint4 *)(gbuf + (uint64_t)gl * NB * P.gcap);
        const uint4 *stage4 = (const uint4 *)stage;
#pragma unroll 4
        for (uint32_t idx = threadIdx.x; idx < NB * (PT_CAP / 4); idx += PT_THREADS) {
            const uint32_t b = idx / (PT_CAP / 4), q = idx % (PT_CAP / 4);
            const uint32_t n = cnt[b];
            if (q * 4 < n) {
                uint4 v = stage4[idx];
                const uint32_t first = stage[b * PT_CAP];
                const uint32_t have = n - q * 4;  // valid entries in this quad (>= 1)
                if (have < 2) v.y = first;
                if (have < 3) v.z = first;
                if (have < 4) v.w = first;
                const uint32_t o = base[b];
                if (o != PT_SKIP) {
                    gb4[(o >> 2) + q] = v;
                } else {
                    const uint64_t fb = g.filter_word_off * 64;
                    overflow_push_cold(ovf_count, ovf, P.ovf_cap, fb + v.x);
                    overflow_push_cold(ovf_count, ovf, P.ovf_cap, fb + v.y);
                    overflow_push_cold(ovf_count, ovf, P.ovf_cap, fb + v.z);
                    overflow_push_cold(ovf_count, ovf, P.ovf_cap, fb + v.w);
                }
            }
        }
        __syncthreads();
    }
}

// Work item = one bucket of one filter = TPB consecutive tiles.  The bucket's positions are pulled
// into registers once (and prefetched for the next bucket while the last store of the current one
// drains); each tile is then zeroed in shared memory, filled with smem atomicOr and written to HBM
// with a single TMA bulk store whose completion is only awaited when the buffer is needed again.
constexpr int TB_REG_ENTRIES = 24;  // positions per thread kept in registers (x256 threads = 6144 per bucket)

__global__ void __launch_bounds__(TB_THREADS, 3)
tile_build_kernel(const DevGenome *__restrict__ genomes, TiledParams P, const uint32_t *__restrict__ gcnt,
                  const uint32_t *__restrict__ gbuf, uint64_t *__restrict__ filters) {
    extern __shared__ __align__(128) uint32_t tb_tile[];
    const uint32_t NB = P.buckets_per_filter;
    const uint32_t n_items = P.n_group * NB;
    const uint32_t tpb_log2 = P.bucket_shift - TB_TILE_LOG2;
    const uint32_t saddr = (uint32_t)__cvta_generic_to_shared(tb_tile);
    uint32_t e[TB_REG_ENTRIES];
    uint32_t n = 0;
    const uint32_t *src = gbuf;
    bool store_pending = false;

    uint32_t item = blockIdx.x;
    if (item < n_items) {
        n = min(gcnt[item], P.gcap);
        src = gbuf + (uint64_t)item * P.gcap;
#pragma unroll
        for (int u = 0; u < TB_REG_ENTRIES; u++) {
            const uint32_t i = u * TB_THREADS + threadIdx.x;
            e[u] = (i < n) ? __ldg(src + i) : 0u;
        }
    }
    for (; item < n_items; item += gridDim.x) {
        const uint32_t gl = item / NB, b = item % NB;
        const uint64_t fw = genomes[P.g0 + gl].filter_word_off;
        for (uint32_t sub = 0; sub < (1u << tpb_log2); sub++) {
            const uint32_t tf = (b << tpb_log2) + sub;
            if (tf >= P.tiles_per_filter) break;
            if (store_pending && threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            __syncthreads();
            uint4 *tv = (uint4 *)tb_tile;
            for (uint32_t i = threadIdx.x; i < TB_TILE_WORDS32 / 4; i += TB_THREADS) tv[i] = make_uint4(0, 0, 0, 0);
            __syncthreads();
#pragma unroll
            for (int u = 0; u < TB_REG_ENTRIES; u++) {
                const uint32_t i = u * TB_THREADS + threadIdx.x;
                if (i < n && (e[u] >> TB_TILE_LOG2) == tf)
                    atomicOr(tb_tile + ((e[u] & ((1u << TB_TILE_LOG2) - 1)) >> 5), 1u << (e[u] & 31));
            }
            for (uint32_t i = TB_REG_ENTRIES * TB_THREADS + threadIdx.x; i < n; i += TB_THREADS) {  // oversized buckets
                const uint32_t p = __ldg(src + i);
                if ((p >> TB_TILE_LOG2) == tf) atomicOr(tb_tile + ((p & ((1u << TB_TILE_LOG2) - 1)) >> 5), 1u << (p & 31));
            }
            // generic-proxy writes to smem must be visible to the async proxy before the bulk store reads them
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncthreads();
            const uint64_t w0 = (uint64_t)tf * (TB_TILE_WORDS32 / 2);  // first u64 word of this tile in its filter
            const uint64_t nw = min((uint64_t)(TB_TILE_WORDS32 / 2), P.words - w0);
            uint64_t *dst = filters + fw + w0;
            const uint32_t bytes = (uint32_t)(nw * 8);
            if (((bytes & 15) == 0) && ((((uintptr_t)dst) & 15) == 0)) {
                if (threadIdx.x == 0) {
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(saddr), "r"(bytes) : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
                store_pending = true;
            } else {
                const uint64_t *tw = (const uint64_t *)tb_tile;
                for (uint64_t i = threadIdx.x; i < nw; i += TB_THREADS) dst[i] = tw[i];
            }
        }
        const uint32_t next = item + gridDim.x;
        if (next < n_items) {  // prefetch the next bucket while the last store drains
            n = min(gcnt[next], P.gcap);
            src = gbuf + (uint64_t)next * P.gcap;
#pragma unroll
            for (int u = 0; u < TB_REG_ENTRIES; u++) {
                const uint32_t i = u * TB_THREADS + threadIdx.x;
                e[u] = (i < n) ? __ldg(src + i) : 0u;
            }
        }
    }
    if (store_pending && threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    __syncthreads();
}

__global__ void __launch_bounds__(256)
overflow_fixup_kernel(const unsigned long long *__restrict__ ovf_count, const uint64_t *__restrict__ ovf, uint64_t cap,
                      uint32_t *__restrict__ filters32) {
    const uint64_t n = min((uint64_t)*ovf_count, cap);
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t a = ovf[i];
        atomicOr(filters32 + (a >> 5), 1u << (a & 31));
    }
}

#include "msbt_fused.cuh"
#include "msbt_phased.cuh"

#ifdef MSBT_FU_TRACE
extern "C" int msbt_debug_trace(unsigned long long *out24, int reset) {
    if (cudaMemcpyFromSymbol(out24, fu_trace, sizeof(unsigned long long) * 24) != cudaSuccess) return -1;
    if (reset) { unsigned long long z[24] = {0}; cudaMemcpyToSymbol(fu_trace, z, sizeof(z)); }
    return 0;
}
#endif

// ---- exact multiplicity path (min_abund >= 2, k <= 32): open-addressing count table.
constexpr uint64_t HT_EMPTY = ~0ULL;  // never a canonical k-mer for k <= 32 (all-T's rc is all-A = 0)

__device__ __forceinline__ uint64_t ht_slot(uint64_t key, uint64_t mask) { return msbt_fmix64(key ^ 0x9E3779B97F4A7C15ULL) & mask; }

__global__ void __launch_bounds__(TILE_THREADS, 4)
kmer_count_insert_kernel(const uint64_t *__restrict__ packed, const uint32_t *__restrict__ run_ends, DevGenome g,
                         uint32_t k, unsigned long long *__restrict__ keys, uint32_t *__restrict__ counts,
                         uint64_t slot_mask) {
    const uint32_t n_words = (g.n_bases + 31) / 32;
    for (uint32_t w = blockIdx.x * TILE_THREADS + threadIdx.x; w < n_words; w += gridDim.x * TILE_THREADS) {
        const uint32_t s0 = w * 32;
        const uint32_t vm = valid_mask(run_ends + g.run_off, g.n_runs, s0, k);
        if (vm == 0) continue;
        KmerWalker<1> kw;
        kw.init(packed + g.packed_word_off + w, k);
        for (uint32_t j = 0; j < 32; j++) {
            if ((vm >> j) & 1) {
                uint64_t lo, hi;
                kw.canonical(lo, hi);
                uint64_t slot = ht_slot(lo, slot_mask);
                while (true) {
                    unsigned long long old = atomicCAS(keys + slot, (unsigned long long)HT_EMPTY, (unsigned long long)lo);
                    if (old == HT_EMPTY || old == lo) {
                        atomicAdd(counts + slot, 1u);
                        break;
                    }
                    slot = (slot + 1) & slot_mask;
                }
            }
            kw.roll(j);
        }
    }
}

__global__ void __launch_bounds__(256)
kmer_count_emit_kernel(const unsigned long long *__restrict__ keys, const uint32_t *__restrict__ counts,
                       uint64_t n_slots, uint32_t min_abund, KmerSpec spec, uint32_t *__restrict__ out) {
    for (uint64_t s = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; s < n_slots; s += (uint64_t)gridDim.x * blockDim.x) {
        unsigned long long key = keys[s];
        if (key != HT_EMPTY && counts[s] >= min_abund) {
            uint64_t pos = msbt_position(msbt_kmer_hash(key, 0, spec.k, spec.pos.seed), spec.pos);
            if (pos < spec.pos.num_bits) atomicOr(out + (pos >> 5), 1u << (pos & 31));
        }
    }
}

// ---- exact multiplicity, compact form (k <= 31, min_abund <= 3): the canonical k-mer occupies 62 bits, so its saturating
// count (1..3) rides in the two top bits of the same 64-bit slot: one table of 8 bytes per slot at load factor <= 2/3
// (64 MiB for a 5 Mbp genome: L2-resident), 0 = empty (an occupied slot always has a non-zero count), one CAS for a
// new k-mer and one more for a repeat.  Keys come from the same unrolled KmerBlock extraction as the other kernels.
constexpr uint64_t HT2_KEY_MASK = (1ull << 62) - 1;

__device__ __forceinline__ void ht2_insert(unsigned long long *__restrict__ tab, uint64_t mask, uint64_t key) {
    uint64_t slot = ht_slot(key, mask);
    const unsigned long long fresh = key | (1ull << 62);
    while (true) {
        unsigned long long old = atomicCAS(tab + slot, 0ull, fresh);
        if (old == 0ull) return;
        if ((old & HT2_KEY_MASK) == key) {
            while ((old >> 62) < 3ull) {  // saturating increment
                const unsigned long long prev = atomicCAS(tab + slot, old, old + (1ull << 62));
                if (prev == old) return;
                old = prev;
            }
            return;
        }
        slot = (slot + 1) & mask;
    }
}

#define MSBT_HT2_STEP(J)                                              \
    {                                                                 \
        uint64_t lo_, hi_;                                            \
        kb.template get<J>(lo_, hi_);                                 \
        if ((vg >> J) & 1u) ht2_insert(tab, slot_mask, lo_);          \
    }

__global__ void __launch_bounds__(TILE_THREADS, 4)
kmer_count2_insert_kernel(const uint64_t *__restrict__ packed, const uint32_t *__restrict__ run_ends, DevGenome g,
                          uint32_t k, unsigned long long *__restrict__ tab, uint64_t slot_mask) {
    const uint32_t n_words = (g.n_bases + 31) / 32;
    for (uint32_t w = blockIdx.x * TILE_THREADS + threadIdx.x; w < n_words; w += gridDim.x * TILE_THREADS) {
        const uint32_t vm = valid_mask(run_ends + g.run_off, g.n_runs, w * 32, k);
        if (vm == 0) continue;
        KmerBlock<1> kb;
        kb.init(packed + g.packed_word_off + w, k);
#pragma unroll 1
        for (uint32_t grp = 0; grp < 4; grp++) {
            const uint32_t vg = (vm >> (8 * grp)) & 0xFFu;
            MSBT_HT2_STEP(0) MSBT_HT2_STEP(1) MSBT_HT2_STEP(2) MSBT_HT2_STEP(3)
            MSBT_HT2_STEP(4) MSBT_HT2_STEP(5) MSBT_HT2_STEP(6) MSBT_HT2_STEP(7)
            kb.advance8();
        }
    }
}

__global__ void __launch_bounds__(256)
kmer_count2_emit_kernel(const unsigned long long *__restrict__ tab, uint64_t n_slots, uint32_t min_abund, KmerSpec spec,
                        uint32_t *__restrict__ out) {
    for (uint64_t s = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; s < n_slots; s += (uint64_t)gridDim.x * blockDim.x) {
        const unsigned long long w = tab[s];
        if (w != 0ull && (w >> 62) >= min_abund) {
            const uint64_t pos = msbt_position(msbt_kmer_hash(w & HT2_KEY_MASK, 0, spec.k, spec.pos.seed), spec.pos);
            if (pos < spec.pos.num_bits) atomicOr(out + (pos >> 5), 1u << (pos & 31));
        }
    }
}

// ---- exact multiplicity, prefiltered form (k <= 31, any min_abund >= 2).  On assemblies almost every k-mer occurs
// once, so an exact counter for all of them (one 64-bit CAS each, above) is wasted work.  Instead:
//   1 mark    every k-mer sets one bit of a one-hash bit table (32 bits per base: ~2 % already-set hits); a k-mer that
//             finds its bit set MAY be a repeat and goes to the candidate list.  Every k-mer that occurs >= 2 times is a
//             candidate (its second occurrence finds the bit its first one set); false positives are k-mers of count 1;
//   2 prepare size the exact table from the number of candidates (device side, no host round trip) and zero it;
//   3 insert  candidates -> keys of the exact table (deduplicated by CAS);
//   4 count   every k-mer looks itself up (a read of a table that is small and cache resident) and, if present, adds 1;
//   5 emit    keys with count >= min_abund set their filter bit.
struct PreState {
    uint32_t n_cand;     // candidates appended by the mark pass
    uint32_t slot_mask;  // exact table size - 1 (prepare)
};
constexpr unsigned long long PRE_OCC = 1ull << 63;  // k <= 31: keys < 2^62, so an occupied slot is never 0
constexpr uint32_t PRE_BLOCK_LIST = 1024;           // candidates a block gathers in shared memory before one global reservation
// The count pass reads a table that holds ~2 % of the k-mers: a two-probe Bloom bitmap of the candidate keys (2^20 bits =
// 128 KiB, filled by the insert pass, copied into shared memory by every CTA of the count pass) answers "not a candidate"
// for ~95 % of the k-mers without an L2 access.
constexpr uint32_t PRE_CB_LOG2_MAX = 20;
constexpr uint32_t PRE_CB_WORDS_MAX = 1u << (PRE_CB_LOG2_MAX - 5);
constexpr int PRE_COUNT_THREADS = 1024;
// the two bitmap probes of a key; `mult`: one multiplication instead of fmix64 (tuning switch)
__device__ __forceinline__ void pre_cb_bits(uint64_t key, uint32_t cb_mask, uint32_t mult, uint32_t &b1, uint32_t &b2) {
    const uint64_t h = mult ? (key ^ (key >> 29)) * 0xC2B2AE3D27D4EB4FULL : msbt_fmix64(key ^ 0xC2B2AE3D27D4EB4FULL);
    b1 = (uint32_t)(h >> 40) & cb_mask;
    b2 = (uint32_t)(h >> 18) & cb_mask;
}

__device__ __forceinline__ void pre_mark(uint64_t key, uint32_t *__restrict__ bits, uint64_t bit_mask, uint64_t *s_list,
                                         uint32_t *s_n, uint64_t *__restrict__ cand, PreState *st) {
    const uint64_t h = msbt_fmix64(key ^ 0xD6E8FEB86659FD93ULL) & bit_mask;
    const uint32_t m = 1u << (h & 31);
    if (atomicOr(bits + (h >> 5), m) & m) {
        const uint32_t i = atomicAdd(s_n, 1u);
        if (i < PRE_BLOCK_LIST) s_list[i] = key;
        else cand[atomicAdd(&st->n_cand, 1u)] = key;  // a very repetitive stretch: straight to the global list
    }
}

#define MSBT_PRE_MARK_STEP(J)                                                        \
    {                                                                                \
        uint64_t lo_, hi_;                                                           \
        kb.template get<J>(lo_, hi_);                                                \
        if ((vg >> J) & 1u) pre_mark(lo_, bits, bit_mask, s_list, &s_n, cand, st);   \
    }

__global__ void __launch_bounds__(TILE_THREADS, 4)
kmer_pre_mark_kernel(const uint64_t *__restrict__ packed, const uint32_t *__restrict__ run_ends, DevGenome g, uint32_t k,
                     uint32_t *__restrict__ bits, uint64_t bit_mask, uint64_t *__restrict__ cand, PreState *st) {
    __shared__ uint64_t s_list[PRE_BLOCK_LIST];
    __shared__ uint32_t s_n, s_base;
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    const uint32_t n_words = (g.n_bases + 31) / 32;
    for (uint32_t w = blockIdx.x * TILE_THREADS + threadIdx.x; w < n_words; w += gridDim.x * TILE_THREADS) {
        const uint32_t vm = valid_mask(run_ends + g.run_off, g.n_runs, w * 32, k);
        if (vm == 0) continue;
        KmerBlock<1> kb;
        kb.init(packed + g.packed_word_off + w, k);
#pragma unroll 1
        for (uint32_t grp = 0; grp < 4; grp++) {
            const uint32_t vg = (vm >> (8 * grp)) & 0xFFu;
            MSBT_PRE_MARK_STEP(0) MSBT_PRE_MARK_STEP(1) MSBT_PRE_MARK_STEP(2) MSBT_PRE_MARK_STEP(3)
            MSBT_PRE_MARK_STEP(4) MSBT_PRE_MARK_STEP(5) MSBT_PRE_MARK_STEP(6) MSBT_PRE_MARK_STEP(7)
            kb.advance8();
        }
    }
    __syncthreads();
    const uint32_t n = min(s_n, PRE_BLOCK_LIST);
    if (threadIdx.x == 0 && n) s_base = atomicAdd(&st->n_cand, n);
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < n; i += TILE_THREADS) cand[s_base + i] = s_list[i];
}

__device__ __forceinline__ uint32_t pre_slots(uint32_t n_cand, uint32_t max_slots) {
    uint32_t slots = 1024;
    while (slots < 2u * n_cand && slots < max_slots) slots <<= 1;
    return slots;
}

__global__ void __launch_bounds__(256)
kmer_pre_prepare_kernel(PreState *st, unsigned long long *__restrict__ tab, uint32_t *__restrict__ cnt, uint32_t max_slots,
                        uint32_t *__restrict__ cbits, uint32_t cb_words) {
    const uint32_t slots = pre_slots(st->n_cand, max_slots);
    for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < cb_words; s += gridDim.x * blockDim.x) cbits[s] = 0u;
    for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < slots; s += gridDim.x * blockDim.x) {
        tab[s] = 0ull;
        cnt[s] = 0u;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) st->slot_mask = slots - 1;  // read by the next kernels only
}

__global__ void __launch_bounds__(256)
kmer_pre_insert_kernel(const PreState *st, const uint64_t *__restrict__ cand, unsigned long long *__restrict__ tab,
                       uint32_t *__restrict__ cbits, uint32_t cb_mask, uint32_t cb_mult) {
    const uint32_t n = st->n_cand, mask = st->slot_mask;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const unsigned long long want = cand[i] | PRE_OCC;
        uint32_t slot = (uint32_t)ht_slot(cand[i], mask);
        while (true) {
            const unsigned long long old = atomicCAS(tab + slot, 0ull, want);
            if (old == 0ull) {  // the thread that placed the key also publishes it in the bitmap
                uint32_t b1, b2;
                pre_cb_bits(cand[i], cb_mask, cb_mult, b1, b2);
                atomicOr(cbits + (b1 >> 5), 1u << (b1 & 31));
                atomicOr(cbits + (b2 >> 5), 1u << (b2 & 31));
                break;
            }
            if (old == want) break;
            slot = (slot + 1) & mask;
        }
    }
}

#define MSBT_PRE_COUNT_STEP(J)                                                       \
    {                                                                                \
        uint64_t lo_, hi_;                                                           \
        kb.template get<J>(lo_, hi_);                                                \
        if ((vg >> J) & 1u) {                                                        \
            const unsigned long long want_ = lo_ | PRE_OCC;                          \
            uint32_t slot_ = (uint32_t)ht_slot(lo_, mask);                           \
            while (true) {                                                           \
                const unsigned long long v_ = __ldg(tab + slot_);                    \
                if (v_ == 0ull) break;                                               \
                if (v_ == want_) { atomicAdd(cnt + slot_, 1u); break; }              \
                slot_ = (slot_ + 1) & mask;                                          \
            }                                                                        \
        }                                                                            \
    }

__global__ void __launch_bounds__(TILE_THREADS, 4)
kmer_pre_count_kernel(const uint64_t *__restrict__ packed, const uint32_t *__restrict__ run_ends, DevGenome g, uint32_t k,
                      const PreState *st, const unsigned long long *__restrict__ tab, uint32_t *__restrict__ cnt) {
    if (st->n_cand == 0) return;  // no k-mer can occur twice
    const uint32_t mask = st->slot_mask;
    const uint32_t n_words = (g.n_bases + 31) / 32;
    for (uint32_t w = blockIdx.x * TILE_THREADS + threadIdx.x; w < n_words; w += gridDim.x * TILE_THREADS) {
        const uint32_t vm = valid_mask(run_ends + g.run_off, g.n_runs, w * 32, k);
        if (vm == 0) continue;
        KmerBlock<1> kb;
        kb.init(packed + g.packed_word_off + w, k);
#pragma unroll 1
        for (uint32_t grp = 0; grp < 4; grp++) {
            const uint32_t vg = (vm >> (8 * grp)) & 0xFFu;
            MSBT_PRE_COUNT_STEP(0) MSBT_PRE_COUNT_STEP(1) MSBT_PRE_COUNT_STEP(2) MSBT_PRE_COUNT_STEP(3)
            MSBT_PRE_COUNT_STEP(4) MSBT_PRE_COUNT_STEP(5) MSBT_PRE_COUNT_STEP(6) MSBT_PRE_COUNT_STEP(7)
            kb.advance8();
        }
    }
}

// The same count with the candidate bitmap in shared memory.  Work unit = a quarter word (8 k-mers): with one 1,024-thread
// CTA per SM a 5 Mbp genome has ~4 units per thread, so the last, partly filled round costs little.
#define MSBT_PRE_COUNT_STEP_CB(J)                                                                      \
    {                                                                                                  \
        uint64_t lo_, hi_;                                                                             \
        kb.template get<J>(lo_, hi_);                                                                  \
        uint32_t b1_, b2_;                                                                             \
        pre_cb_bits(lo_, cb_mask, cb_mult, b1_, b2_);                                                  \
        const uint32_t hit_ = (s_cb[b1_ >> 5] >> (b1_ & 31)) & (s_cb[b2_ >> 5] >> (b2_ & 31)) & (vg >> J) & 1u; \
        if (hit_) {                                                                                    \
            const unsigned long long want_ = lo_ | PRE_OCC;                                            \
            uint32_t slot_ = (uint32_t)ht_slot(lo_, mask);                                             \
            while (true) {                                                                             \
                const unsigned long long v_ = __ldg(tab + slot_);                                      \
                if (v_ == 0ull) break;                                                                 \
                if (v_ == want_) { atomicAdd(cnt + slot_, 1u); break; }                                \
                slot_ = (slot_ + 1) & mask;                                                            \
            }                                                                                          \
        }                                                                                              \
    }

// `ug` = groups of 8 k-mers per work unit (1, 2 or 4): with one or two 1,024-thread CTAs per SM a 5 Mbp genome has only a
// few units per thread, so finer units make the last, partly filled round cheap at the price of repeated set-up.
__global__ void __launch_bounds__(PRE_COUNT_THREADS, 2)
kmer_pre_count_cb_kernel(const uint64_t *__restrict__ packed, const uint32_t *__restrict__ run_ends, DevGenome g, uint32_t k,
                         const PreState *st, const unsigned long long *__restrict__ tab, uint32_t *__restrict__ cnt,
                         const uint32_t *__restrict__ cbits, uint32_t cb_mask, uint32_t cb_mult, uint32_t ug) {
    extern __shared__ __align__(16) uint32_t s_cb[];
    if (st->n_cand == 0) return;  // no k-mer can occur twice
    for (uint32_t i = threadIdx.x; i < (cb_mask + 1u) / 128u; i += PRE_COUNT_THREADS) ((uint4 *)s_cb)[i] = __ldcg((const uint4 *)cbits + i);
    __syncthreads();
    const uint32_t mask = st->slot_mask;
    const uint32_t upw = 4u / ug;  // units per packed word
    const uint32_t n_units = ((g.n_bases + 31) / 32) * upw;
    for (uint32_t u = blockIdx.x * PRE_COUNT_THREADS + threadIdx.x; u < n_units; u += gridDim.x * PRE_COUNT_THREADS) {
        const uint32_t w = u / upw, g0 = (u % upw) * ug;
        const uint32_t vm = valid_mask(run_ends + g.run_off, g.n_runs, w * 32, k) >> (8 * g0);
        if ((ug == 4 ? vm : (vm & ((1u << (8 * ug)) - 1u))) == 0) continue;
        KmerBlock<1> kb;
        kb.init(packed + g.packed_word_off + w, k);
        for (uint32_t a = 0; a < g0; a++) kb.advance8();
#pragma unroll 1
        for (uint32_t grp = 0; grp < ug; grp++) {
            const uint32_t vg = (vm >> (8 * grp)) & 0xFFu;
            MSBT_PRE_COUNT_STEP_CB(0) MSBT_PRE_COUNT_STEP_CB(1) MSBT_PRE_COUNT_STEP_CB(2) MSBT_PRE_COUNT_STEP_CB(3)
            MSBT_PRE_COUNT_STEP_CB(4) MSBT_PRE_COUNT_STEP_CB(5) MSBT_PRE_COUNT_STEP_CB(6) MSBT_PRE_COUNT_STEP_CB(7)
            kb.advance8();
        }
    }
}

// zero-fill with streaming stores (evict-first in L2): a filter zeroed next to the mark pass of another genome must not push
// that genome's bit table out of L2
__global__ void __launch_bounds__(256)
bf_zero_stream_kernel(uint4 *__restrict__ dst, uint64_t vec) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < vec; i += (uint64_t)gridDim.x * blockDim.x)
        __stcs(dst + i, make_uint4(0, 0, 0, 0));
}

__global__ void __launch_bounds__(256)
kmer_pre_emit_kernel(const PreState *st, const unsigned long long *__restrict__ tab, const uint32_t *__restrict__ cnt,
                     uint32_t min_abund, KmerSpec spec, uint32_t *__restrict__ out) {
    if (st->n_cand == 0) return;
    const uint32_t slots = st->slot_mask + 1;
    for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < slots; s += gridDim.x * blockDim.x) {
        const unsigned long long v = tab[s];
        if (v != 0ull && cnt[s] >= min_abund) {
            const uint64_t pos = msbt_position(msbt_kmer_hash(v & ~PRE_OCC, 0, spec.k, spec.pos.seed), spec.pos);
            if (pos < spec.pos.num_bits) atomicOr(out + (pos >> 5), 1u << (pos & 31));
        }
    }
}

static uint32_t ceil_log2_u64(uint64_t x) {
    uint32_t l = 0;
    while ((1ull << l) < x && l < 63) l++;
    return l;
}

static int stage_genome_table(msbt_ctx *ctx, const std::vector<DevGenome> &dg, size_t *table_bytes, cudaStream_t stream) {
    const size_t table_b = align_up(sizeof(DevGenome) * dg.size(), 256);
    int rc = ensure_scratch(ctx, table_b);
    if (rc) return rc;
    rc = stage_begin(ctx, table_b);
    if (rc) return rc;
    rc = stage_to_device(ctx, ctx->d_scratch, dg.data(), sizeof(DevGenome) * dg.size(), 0, stream);
    if (rc) return rc;
    rc = stage_end(ctx, stream);
    if (rc) return rc;
    *table_bytes = table_b;
    return MSBT_OK;
}

static int build_tiled(msbt_ctx *ctx, const uint64_t *d_packed, const uint32_t *d_run_ends,
                       const std::vector<DevGenome> &dg, uint64_t total_tiles, const KmerSpec &spec, uint64_t words,
                       uint64_t *d_filters, cudaStream_t stream) {
    const uint32_t n_genomes = (uint32_t)dg.size();
    const uint64_t num_bits = spec.pos.num_bits;
    const uint32_t k = spec.k;
    // bucket geometry
    uint32_t bucket_shift = TB_TILE_LOG2;
    while (((num_bits + (1ull << bucket_shift) - 1) >> bucket_shift) > PT_MAX_BUCKETS) bucket_shift++;
    const uint32_t NB = (uint32_t)((num_bits + (1ull << bucket_shift) - 1) >> bucket_shift);
    const uint32_t TPF = (uint32_t)((num_bits + (1ull << TB_TILE_LOG2) - 1) >> TB_TILE_LOG2);
    (void)ceil_log2_u64;

    auto kmers_of = [&](const DevGenome &g) -> uint64_t { return g.n_bases >= k ? (uint64_t)g.n_bases - k + 1 : 0; };
    auto gcap_for = [&](uint64_t max_kmers) -> uint64_t {
        // positions are uniform over [0, modulus): expected entries per bucket, plus 8 sigma and slack
        const double frac = std::min(1.0, (double)(1ull << bucket_shift) / (double)spec.pos.modulus);
        const double e = (double)max_kmers * frac;
        // each chunk that touches a bucket may add up to 3 pad entries (expected 1.5)
        const double chunks = (double)((max_kmers + TILE_BASES - 1) / TILE_BASES);
        const double touched = chunks * std::min(1.0, (double)TILE_BASES * frac);
        uint64_t cap = (uint64_t)(e + 8.0 * sqrt(e) + 2.0 * touched + 64.0);
        cap = std::min<uint64_t>(cap, max_kmers + 3 * (uint64_t)chunks + 8);
        return (cap + 7) / 8 * 8;
    };

    // groups of consecutive genomes whose bucket scratch stays around 64 MB (L2-resident)
    const uint64_t SCRATCH_TARGET = 64ull << 20;
    struct Group { uint32_t g0, n; uint64_t gcap, kmers; };
    std::vector<Group> groups;
    uint64_t max_gbuf = 0, max_kmers_group = 0, max_gcnt = 0;
    for (uint32_t i = 0; i < n_genomes;) {
        uint32_t j = i;
        uint64_t mk = 0, sum_k = 0;
        while (j < n_genomes) {
            const uint64_t mk2 = std::max(mk, kmers_of(dg[j]));
            const uint64_t need = (uint64_t)(j - i + 1) * NB * gcap_for(mk2) * 4;
            if (j > i && need > SCRATCH_TARGET) break;
            if ((uint64_t)(j - i + 1) * TPF > (1u << 30)) break;
            mk = mk2;
            sum_k += kmers_of(dg[j]);
            j++;
        }
        Group gr{i, j - i, gcap_for(mk), sum_k};
        if (((uint64_t)NB * gr.gcap) >> 32) return 1;  // not applicable: caller falls back to the direct variant
        groups.push_back(gr);
        max_gbuf = std::max(max_gbuf, (uint64_t)gr.n * NB * gr.gcap * 4);
        max_gcnt = std::max(max_gcnt, (uint64_t)gr.n * NB * 4);
        max_kmers_group = std::max(max_kmers_group, sum_k);
        i = j;
    }

    size_t table_b = 0;
    // scratch layout: [genome table][ovf_count (256 B)][gcnt][gbuf][ovf list]
    const size_t t_b = align_up(sizeof(DevGenome) * n_genomes, 256);
    const size_t cnt_off = t_b, gcnt_off = cnt_off + 256;
    const size_t gbuf_off = align_up(gcnt_off + max_gcnt, 256);
    const size_t ovf_off = align_up(gbuf_off + max_gbuf, 256);
    const size_t total = ovf_off + max_kmers_group * 8 + 256;
    int rc = ensure_scratch(ctx, total);
    if (rc) return rc;
    rc = stage_genome_table(ctx, dg, &table_b, stream);
    if (rc) return rc;
    char *S = (char *)ctx->d_scratch;
    const DevGenome *d_genomes = (const DevGenome *)S;
    unsigned long long *ovf_count = (unsigned long long *)(S + cnt_off);
    uint32_t *gcnt = (uint32_t *)(S + gcnt_off);
    uint32_t *gbuf = (uint32_t *)(S + gbuf_off);
    uint64_t *ovf = (uint64_t *)(S + ovf_off);

    const size_t pt_smem = PT_SMEM_BYTES;
    static_assert(PT_CAP == 16, "PT_SMEM_BYTES assumes 16 staging slots");
    const size_t tb_smem = (size_t)TB_TILE_WORDS32 * 4;
    if (!ctx->tiled_attr_set) {
        const int pt_max = (int)PT_SMEM_BYTES;
#define MSBT_PT_ATTR(KW, HV) \
        CU_TRY(ctx, cudaFuncSetAttribute(kmer_partition_kernel<KW, HV, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, pt_max)); \
        CU_TRY(ctx, cudaFuncSetAttribute(kmer_partition_kernel<KW, HV, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, pt_max)); \
        CU_TRY(ctx, cudaFuncSetAttribute(kmer_partition_kernel<KW, HV, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, pt_max)); \
        CU_TRY(ctx, cudaFuncSetAttribute(kmer_partition_kernel<KW, HV, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, pt_max));
        MSBT_PT_ATTR(1, 0) MSBT_PT_ATTR(2, 1) MSBT_PT_ATTR(2, 2)
#undef MSBT_PT_ATTR
        CU_TRY(ctx, cudaFuncSetAttribute(tile_build_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tb_smem));
        ctx->tiled_attr_set = true;
    }
    (void)total_tiles;
    for (const Group &gr : groups) {
        TiledParams P;
        P.g0 = gr.g0;
        P.n_group = gr.n;
        P.tile0 = dg[gr.g0].tile_begin;
        const DevGenome &last = dg[gr.g0 + gr.n - 1];
        P.n_chunks = last.tile_begin + (uint32_t)(((uint64_t)last.n_bases + TILE_BASES - 1) / TILE_BASES) - P.tile0;
        P.bucket_shift = bucket_shift;
        P.buckets_per_filter = NB;
        P.tiles_per_filter = TPF;
        P.gcap = (uint32_t)gr.gcap;
        P.words = words;
        P.ovf_cap = max_kmers_group;
        // ovf_count and gcnt are adjacent: one memset clears both
        CU_TRY(ctx, cudaMemsetAsync(S + cnt_off, 0, 256 + (size_t)gr.n * NB * 4, stream));
        if (P.n_chunks) {
            const int grid = (int)std::min<uint64_t>(P.n_chunks, (uint64_t)ctx->num_sms * 3);
#define MSBT_PT_LAUNCH(KW, HV)                                                                                         \
    {                                                                                                                  \
        if (p2 && full) kmer_partition_kernel<KW, HV, true, true><<<grid, PT_THREADS, pt_smem, stream>>>(d_packed, d_run_ends, d_genomes, P, spec, gcnt, gbuf, ovf_count, ovf); \
        else if (p2) kmer_partition_kernel<KW, HV, true, false><<<grid, PT_THREADS, pt_smem, stream>>>(d_packed, d_run_ends, d_genomes, P, spec, gcnt, gbuf, ovf_count, ovf); \
        else if (full) kmer_partition_kernel<KW, HV, false, true><<<grid, PT_THREADS, pt_smem, stream>>>(d_packed, d_run_ends, d_genomes, P, spec, gcnt, gbuf, ovf_count, ovf); \
        else kmer_partition_kernel<KW, HV, false, false><<<grid, PT_THREADS, pt_smem, stream>>>(d_packed, d_run_ends, d_genomes, P, spec, gcnt, gbuf, ovf_count, ovf); \
    }
            const int hv = msbt_hash_variant(k);
            const bool p2 = spec.pos.pow2 != 0, full = spec.pos.modulus <= spec.pos.num_bits;
            if (hv == 0) MSBT_PT_LAUNCH(1, 0)
            else if (hv == 1) MSBT_PT_LAUNCH(2, 1)
            else MSBT_PT_LAUNCH(2, 2)
#undef MSBT_PT_LAUNCH
            LAUNCH_CHECK(ctx);
        }
        {
            const uint64_t n_items = (uint64_t)gr.n * NB;
            const int grid = (int)std::min<uint64_t>(n_items, (uint64_t)ctx->num_sms * 3);
            tile_build_kernel<<<grid, TB_THREADS, tb_smem, stream>>>(d_genomes, P, gcnt, gbuf, d_filters);
            LAUNCH_CHECK(ctx);
        }
        if (P.n_chunks) {
            overflow_fixup_kernel<<<ctx->num_sms, 256, 0, stream>>>(ovf_count, ovf, P.ovf_cap, (uint32_t *)d_filters);
            LAUNCH_CHECK(ctx);
        }
    }
    return MSBT_OK;
}


static int launch_direct(msbt_ctx *ctx, const uint64_t *d_packed, const uint32_t *d_run_ends, const DevGenome *d_genomes,
                         uint32_t n_genomes, uint64_t tiles, const KmerSpec &spec, uint64_t *d_filters,
                         const uint32_t *only_dirty, cudaStream_t stream, uint64_t tile_base = 0) {
    if (tiles <= tile_base) return MSBT_OK;
    const int grid = (int)std::min<uint64_t>(tiles - tile_base, (uint64_t)ctx->num_sms * 8);
#define MSBT_DIRECT_LAUNCH(KW, HV, P2)                                                                          \
    kmer_bloom_direct_kernel<KW, HV, P2><<<grid, TILE_THREADS, 0, stream>>>(d_packed, d_run_ends, d_genomes, n_genomes, \
                                                                           (uint32_t)tiles, spec, (uint32_t *)d_filters, only_dirty, (uint32_t)tile_base)
    const int hv = msbt_hash_variant(spec.k);
    const bool p2 = spec.pos.pow2 != 0;
    if (hv == 0) { if (p2) MSBT_DIRECT_LAUNCH(1, 0, true); else MSBT_DIRECT_LAUNCH(1, 0, false); }
    else if (hv == 1) { if (p2) MSBT_DIRECT_LAUNCH(2, 1, true); else MSBT_DIRECT_LAUNCH(2, 1, false); }
    else { if (p2) MSBT_DIRECT_LAUNCH(2, 2, true); else MSBT_DIRECT_LAUNCH(2, 2, false); }
#undef MSBT_DIRECT_LAUNCH
    LAUNCH_CHECK(ctx);
    return MSBT_OK;
}

// Variant 3: the fused persistent kernel of msbt_fused.cuh.  Returns 1 if the geometry does not apply.
static int build_fused(msbt_ctx *ctx, const uint64_t *d_packed, const uint32_t *d_run_ends, const std::vector<DevGenome> &dg_in,
                       uint64_t tiles, const KmerSpec &spec, uint64_t words, uint64_t *d_filters, cudaStream_t stream) {
    std::vector<DevGenome> dg(dg_in);
    uint64_t fchunks = 0;
    for (DevGenome &g : dg) {
        g.fchunk_begin = (uint32_t)fchunks;
        fchunks += ((uint64_t)g.n_bases + FU_CHUNK_BASES - 1) / FU_CHUNK_BASES;
        if (fchunks >> 31) return 1;
    }
    const uint32_t n_genomes = (uint32_t)dg.size();
    const uint64_t num_bits = spec.pos.num_bits;
    const uint32_t k = spec.k;
    uint32_t shift = 7;  // a tile is at least 128 bits (one 16-byte bulk store)
    // at most 512 buckets below 2^30 bits, 2048 from there (msbt_fused.cuh: FuGeom); MSBT_FU_NBMAX=512|2048 forces one (tuning)
    uint32_t nbm = num_bits < (1ull << 30) ? FU_NBM_SMALL : FU_NBM_LARGE;
    if (const char *e = getenv("MSBT_FU_NBMAX")) { const int v = atoi(e); if (v == (int)FU_NBM_SMALL || v == (int)FU_NBM_LARGE) nbm = (uint32_t)v; }
    const uint32_t stage_words = nbm == FU_NBM_SMALL ? FuGeom<FU_NBM_SMALL>::STAGE_WORDS : FuGeom<FU_NBM_LARGE>::STAGE_WORDS;
    const size_t smem_bytes = nbm == FU_NBM_SMALL ? FuGeom<FU_NBM_SMALL>::SMEM_BYTES : FuGeom<FU_NBM_LARGE>::SMEM_BYTES;
    while (((num_bits + (1ull << shift) - 1) >> shift) > nbm) shift++;
    const uint32_t NB = (uint32_t)((num_bits + (1ull << shift) - 1) >> shift);
    if (fchunks == 0 || (fchunks >> 31) || ((uint64_t)n_genomes * NB >> 28)) return 1;  // counters: 4 B per (genome, bucket)

    FusedParams P;
    memset(&P, 0, sizeof(P));
    P.bucket_shift = shift;
    P.NB = NB;
    P.cap = std::min<uint32_t>((stage_words / (NB + 1)) & ~3u, 252u);  // n is packed into 8 bits by the flush
    P.quads = P.cap / 4;
    P.lanes_per_row = 1;
    while (P.lanes_per_row < P.quads && P.lanes_per_row < 32) P.lanes_per_row <<= 1;
    P.quad_magic = (uint32_t)(((1ull << 32) + P.quads - 1) / P.quads);
    if ((uint64_t)NB * P.quads * P.quads >= (1ull << 32)) return 1;  // exactness bound of the magic division
    P.tile_log2 = std::min(shift, FU_TILE_LOG2);
    P.tiles_per_bucket_log2 = shift - P.tile_log2;
    P.words = words;
    P.n_genomes = n_genomes;
    P.n_chunks_total = (uint32_t)fchunks;
    P.nb_magic = (uint32_t)(((1ull << 32) + NB - 1) / NB);
    // lean B loop: bucket == one full 64 KiB tile, every tile complete, every destination 16-byte aligned
    P.fast_b = (shift >= FU_TILE_LOG2) && (num_bits % (1ull << shift) == 0) && ((((uintptr_t)d_filters) & 15) == 0);  // whole buckets of full tiles
    for (const DevGenome &g : dg) P.fast_b = P.fast_b && ((g.filter_word_off & 1) == 0);
    // exactness of item / NB by multiply-high: item * (magic * NB - 2^32) < 2^32
    if (P.fast_b && (uint64_t)n_genomes * NB * ((uint64_t)P.nb_magic * NB - (1ull << 32)) >= (1ull << 32)) P.fast_b = 0;

    uint64_t max_kmers = 0;
    for (const DevGenome &g : dg) max_kmers = std::max<uint64_t>(max_kmers, g.n_bases >= k ? (uint64_t)g.n_bases - k + 1 : 0);
    {
        // positions are uniform over [0, modulus): expected entries per bucket + 8 sigma; every chunk that touches a
        // bucket may add up to 3 pad entries, a full staging row spills 4 entries per k-mer (rare)
        const double frac = std::min(1.0, (double)(1ull << shift) / (double)spec.pos.modulus);
        const double e = (double)max_kmers * frac;
        const double chunks = (double)((max_kmers + FU_CHUNK_BASES - 1) / FU_CHUNK_BASES);
        const double touched = chunks * std::min(1.0, (double)FU_CHUNK_BASES * frac);
        uint64_t cap = (uint64_t)(1.02 * e + 8.0 * sqrt(e) + 3.0 * touched + 64.0);
        cap = std::min<uint64_t>(cap, max_kmers + 3 * (uint64_t)chunks + 8);
        cap = (cap + 7) / 8 * 8;
        if ((cap * NB) >> 32) return 1;
        P.gcap = (uint32_t)cap;
    }
    const uint64_t slot_bytes = (uint64_t)NB * P.gcap * 4;
    P.ring = (uint32_t)std::max<uint64_t>(2, std::min<uint64_t>(4, (100ull << 20) / std::max<uint64_t>(slot_bytes, 1)));

    // scratch: [genome table][sync 64 B][done][consumed][dirty][gcnt n_genomes x NB] | [gbuf ring x NB x gcap]
    const size_t t_b = align_up(sizeof(DevGenome) * n_genomes, 256);
    const size_t sync_off = t_b;
    const size_t done_off = sync_off + 64;
    const size_t cons_off = done_off + (size_t)n_genomes * 4;
    const size_t dirty_off = cons_off + (size_t)n_genomes * 4;
    const size_t gcnt_off = align_up(dirty_off + (size_t)n_genomes * 4, 16);
    const size_t zero_end = gcnt_off + (size_t)n_genomes * NB * 4;
    const size_t gbuf_off = align_up(zero_end, 256);
    const size_t total = gbuf_off + (size_t)P.ring * slot_bytes;
    int rc = ensure_scratch(ctx, total);
    if (rc) return rc == MSBT_ENOMEM ? 1 : rc;
    size_t table_b = 0;
    rc = stage_genome_table(ctx, dg, &table_b, stream);
    if (rc) return rc;
    char *S = (char *)ctx->d_scratch;
    P.packed = d_packed;
    P.run_ends = d_run_ends;
    P.genomes = (const DevGenome *)S;
    P.sync = (uint32_t *)(S + sync_off);
    P.done = (uint32_t *)(S + done_off);
    P.consumed = (uint32_t *)(S + cons_off);
    P.dirty = (uint32_t *)(S + dirty_off);
    P.gcnt = (uint32_t *)(S + gcnt_off);
    P.gbuf = (uint32_t *)(S + gbuf_off);
    P.filters = d_filters;
    CU_TRY(ctx, cudaMemsetAsync(S + sync_off, 0, zero_end - sync_off, stream));

    if (!ctx->fused_attr_set) {
#define MSBT_FU_ATTR1(KW, HV, NBM_)                                                                                                          \
        CU_TRY(ctx, cudaFuncSetAttribute(kmer_bloom_fused_kernel<KW, HV, true, true, NBM_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FuGeom<NBM_>::SMEM_BYTES)); \
        CU_TRY(ctx, cudaFuncSetAttribute(kmer_bloom_fused_kernel<KW, HV, true, false, NBM_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FuGeom<NBM_>::SMEM_BYTES)); \
        CU_TRY(ctx, cudaFuncSetAttribute(kmer_bloom_fused_kernel<KW, HV, false, true, NBM_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FuGeom<NBM_>::SMEM_BYTES)); \
        CU_TRY(ctx, cudaFuncSetAttribute(kmer_bloom_fused_kernel<KW, HV, false, false, NBM_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FuGeom<NBM_>::SMEM_BYTES));
#define MSBT_FU_ATTR(KW, HV) MSBT_FU_ATTR1(KW, HV, FU_NBM_SMALL) MSBT_FU_ATTR1(KW, HV, FU_NBM_LARGE)
        MSBT_FU_ATTR(1, 0) MSBT_FU_ATTR(2, 1) MSBT_FU_ATTR(2, 2)
#undef MSBT_FU_ATTR
#undef MSBT_FU_ATTR1
        ctx->fused_attr_set = true;
    }
    const int grid = ctx->num_sms;
#define MSBT_FU_LAUNCH1(KW, HV, NBM_)                                                                                        \
    {                                                                                                                        \
        if (p2 && full) kmer_bloom_fused_kernel<KW, HV, true, true, NBM_><<<grid, FU_THREADS, smem_bytes, stream>>>(P, spec);    \
        else if (p2) kmer_bloom_fused_kernel<KW, HV, true, false, NBM_><<<grid, FU_THREADS, smem_bytes, stream>>>(P, spec);      \
        else if (full) kmer_bloom_fused_kernel<KW, HV, false, true, NBM_><<<grid, FU_THREADS, smem_bytes, stream>>>(P, spec);    \
        else kmer_bloom_fused_kernel<KW, HV, false, false, NBM_><<<grid, FU_THREADS, smem_bytes, stream>>>(P, spec);             \
    }
#define MSBT_FU_LAUNCH(KW, HV) { if (nbm == FU_NBM_SMALL) MSBT_FU_LAUNCH1(KW, HV, FU_NBM_SMALL) else MSBT_FU_LAUNCH1(KW, HV, FU_NBM_LARGE) }
    const int hv = msbt_hash_variant(k);
    const bool p2 = spec.pos.pow2 != 0, full = spec.pos.modulus <= spec.pos.num_bits;
    if (hv == 0) MSBT_FU_LAUNCH(1, 0)
    else if (hv == 1) MSBT_FU_LAUNCH(2, 1)
    else MSBT_FU_LAUNCH(2, 2)
#undef MSBT_FU_LAUNCH
#undef MSBT_FU_LAUNCH1
    LAUNCH_CHECK(ctx);
    // genomes whose ring buckets overflowed (skewed inputs) are completed by the direct kernel
    return launch_direct(ctx, d_packed, d_run_ends, P.genomes, n_genomes, tiles, spec, d_filters, P.dirty, stream);
}

// Variant 4: the phased persistent kernel of msbt_phased.cuh.  Returns 1 if the geometry does not apply.
static int build_phased(msbt_ctx *ctx, const uint64_t *d_packed, const uint32_t *d_run_ends, const std::vector<DevGenome> &dg_in,
                        uint64_t tiles, const KmerSpec &spec, uint64_t words, uint64_t *d_filters, cudaStream_t stream) {
    std::vector<DevGenome> dg(dg_in);
    uint64_t chunks = 0;
    for (DevGenome &g : dg) {
        g.fchunk_begin = (uint32_t)chunks;
        chunks += ((uint64_t)g.n_bases + PH_CHUNK_BASES - 1) / PH_CHUNK_BASES;
        if (chunks >> 31) return 1;
    }
    const uint32_t n_genomes = (uint32_t)dg.size();
    const uint64_t num_bits = spec.pos.num_bits;
    const uint32_t k = spec.k;
    uint32_t shift = 7;  // a tile is at least 128 bits (one 16-byte bulk store)
    while (((num_bits + (1ull << shift) - 1) >> shift) > PH_MAX_BUCKETS) shift++;
    const uint32_t NB = (uint32_t)((num_bits + (1ull << shift) - 1) >> shift);
    if (chunks == 0 || ((uint64_t)n_genomes * NB >> 28)) return 1;  // counters: 4 B per (genome, bucket)

    PhasedParams P;
    memset(&P, 0, sizeof(P));
    P.bucket_shift = shift;
    P.NB = NB;
    P.cap = (PH_AREA_WORDS / (NB + 1)) & ~3u;
    P.quads = P.cap / 4;
    P.tile_log2 = std::min(shift, PH_TILE_LOG2);
    P.tiles_per_bucket_log2 = shift - P.tile_log2;
    P.words = words;
    P.n_genomes = n_genomes;
    P.n_chunks_total = (uint32_t)chunks;
    P.n_items = n_genomes * NB;

    uint64_t max_kmers = 0;
    for (const DevGenome &g : dg) max_kmers = std::max<uint64_t>(max_kmers, g.n_bases >= k ? (uint64_t)g.n_bases - k + 1 : 0);
    {
        // positions are uniform over [0, modulus): expected entries per bucket + 8 sigma; every chunk that touches a
        // bucket may add up to 3 pad entries, a full staging row spills 4 entries per k-mer (rare)
        const double frac = std::min(1.0, (double)(1ull << shift) / (double)spec.pos.modulus);
        const double e = (double)max_kmers * frac;
        const double nch = (double)((max_kmers + PH_CHUNK_BASES - 1) / PH_CHUNK_BASES);
        const double touched = nch * std::min(1.0, (double)PH_CHUNK_BASES * frac);
        uint64_t cap = (uint64_t)(1.03 * e + 8.0 * sqrt(e) + 3.0 * touched + 64.0);
        cap = std::min<uint64_t>(cap, max_kmers + 3 * (uint64_t)nch + 8);
        cap = (cap + 7) / 8 * 8;
        if ((cap * NB) >> 32) return 1;
        P.gcap = (uint32_t)cap;
    }
    const uint64_t slot_bytes = (uint64_t)NB * P.gcap * 4;
    P.ring = (uint32_t)std::max<uint64_t>(2, std::min<uint64_t>(4, (100ull << 20) / std::max<uint64_t>(slot_bytes, 1)));

    // scratch: [genome table][sync 256 B][done][consumed][dirty][gcnt n_genomes x NB] | [gbuf ring x NB x gcap]
    const size_t t_b = align_up(sizeof(DevGenome) * n_genomes, 256);
    const size_t sync_off = t_b;
    const size_t done_off = sync_off + 256;
    const size_t cons_off = done_off + (size_t)n_genomes * 4;
    const size_t dirty_off = cons_off + (size_t)n_genomes * 4;
    const size_t gcnt_off = align_up(dirty_off + (size_t)n_genomes * 4, 16);
    const size_t zero_end = gcnt_off + (size_t)n_genomes * NB * 4;
    const size_t gbuf_off = align_up(zero_end, 256);
    const size_t total = gbuf_off + (size_t)P.ring * slot_bytes;
    int rc = ensure_scratch(ctx, total);
    if (rc) return rc == MSBT_ENOMEM ? 1 : rc;
    size_t table_b = 0;
    rc = stage_genome_table(ctx, dg, &table_b, stream);
    if (rc) return rc;
    char *S = (char *)ctx->d_scratch;
    P.packed = d_packed;
    P.run_ends = d_run_ends;
    P.genomes = (const DevGenome *)S;
    P.sync = (uint32_t *)(S + sync_off);
    P.done = (uint32_t *)(S + done_off);
    P.consumed = (uint32_t *)(S + cons_off);
    P.dirty = (uint32_t *)(S + dirty_off);
    P.gcnt = (uint32_t *)(S + gcnt_off);
    P.gbuf = (uint32_t *)(S + gbuf_off);
    P.filters = d_filters;
    CU_TRY(ctx, cudaMemsetAsync(S + sync_off, 0, zero_end - sync_off, stream));

    if (!ctx->phased_attr_set) {
#define MSBT_PH_ATTR(KW, HV) \
        CU_TRY(ctx, cudaFuncSetAttribute(kmer_bloom_phased_kernel<KW, HV, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PH_SMEM_BYTES)); \
        CU_TRY(ctx, cudaFuncSetAttribute(kmer_bloom_phased_kernel<KW, HV, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PH_SMEM_BYTES)); \
        CU_TRY(ctx, cudaFuncSetAttribute(kmer_bloom_phased_kernel<KW, HV, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PH_SMEM_BYTES)); \
        CU_TRY(ctx, cudaFuncSetAttribute(kmer_bloom_phased_kernel<KW, HV, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PH_SMEM_BYTES));
        MSBT_PH_ATTR(1, 0) MSBT_PH_ATTR(2, 1) MSBT_PH_ATTR(2, 2)
#undef MSBT_PH_ATTR
        ctx->phased_attr_set = true;
    }
    const int grid = ctx->num_sms;
#define MSBT_PH_LAUNCH(KW, HV)                                                                                     \
    {                                                                                                              \
        if (p2 && full) kmer_bloom_phased_kernel<KW, HV, true, true><<<grid, PH_THREADS, PH_SMEM_BYTES, stream>>>(P, spec);   \
        else if (p2) kmer_bloom_phased_kernel<KW, HV, true, false><<<grid, PH_THREADS, PH_SMEM_BYTES, stream>>>(P, spec);     \
        else if (full) kmer_bloom_phased_kernel<KW, HV, false, true><<<grid, PH_THREADS, PH_SMEM_BYTES, stream>>>(P, spec);   \
        else kmer_bloom_phased_kernel<KW, HV, false, false><<<grid, PH_THREADS, PH_SMEM_BYTES, stream>>>(P, spec);            \
    }
    const int hv = msbt_hash_variant(k);
    const bool p2 = spec.pos.pow2 != 0, full = spec.pos.modulus <= spec.pos.num_bits;
    if (hv == 0) MSBT_PH_LAUNCH(1, 0)
    else if (hv == 1) MSBT_PH_LAUNCH(2, 1)
    else MSBT_PH_LAUNCH(2, 2)
#undef MSBT_PH_LAUNCH
    LAUNCH_CHECK(ctx);
    // genomes whose ring buckets overflowed (skewed inputs) are completed by the direct kernel
    return launch_direct(ctx, d_packed, d_run_ends, P.genomes, n_genomes, tiles, spec, d_filters, P.dirty, stream);
}

extern "C" int msbt_kmer_bloom_build(msbt_ctx *ctx, const uint64_t *d_packed, const uint32_t *d_run_ends,
                                     const msbt_genome_desc *h_descs, uint32_t n_genomes, uint32_t k,
                                     uint64_t seed, uint64_t modulus, uint64_t num_bits, uint32_t min_abund,
                                     uint64_t *d_filters, uint64_t filter_stride_words, int variant,
                                     void *stream_) {
    if (!ctx) return MSBT_EINVAL;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (k < 1 || k > 64) return fail(ctx, MSBT_EINVAL, "k=%u outside 1..64", k);
    if (modulus == 0 || num_bits == 0) return fail(ctx, MSBT_EINVAL, "modulus and num_bits must be > 0");
    const uint64_t words = (num_bits + 63) / 64;
    if (filter_stride_words < words) return fail(ctx, MSBT_EINVAL, "filter stride %llu < %llu words", (unsigned long long)filter_stride_words, (unsigned long long)words);
    if (n_genomes == 0) return MSBT_OK;
    if (!d_packed || !d_run_ends || !h_descs || !d_filters) return fail(ctx, MSBT_EINVAL, "null pointer argument");
    if (min_abund >= 2 && k > 32) return fail(ctx, MSBT_EINVAL, "min_abund >= 2 supports k <= 32 only (k=%u)", k);
    if (variant < 0 || variant > 4) return fail(ctx, MSBT_EINVAL, "unknown variant %d", variant);
    CTX_DEVICE(ctx);

    KmerSpec spec;
    spec.pos = msbt_make_pos_spec(modulus, num_bits, seed);
    spec.k = k;

    // device genome table with tile prefix
    std::vector<DevGenome> dg(n_genomes);
    uint64_t tiles = 0;
    uint64_t max_bases = 0;
    for (uint32_t i = 0; i < n_genomes; i++) {
        const msbt_genome_desc &d = h_descs[i];
        if (d.n_bases >> 32) return fail(ctx, MSBT_EINVAL, "genome %u has >= 2^32 bases", i);
        dg[i].packed_word_off = d.packed_word_off;
        dg[i].run_off = d.run_off;
        dg[i].filter_word_off = (uint64_t)d.filter_index * filter_stride_words;
        dg[i].n_bases = (uint32_t)d.n_bases;
        dg[i].n_runs = d.n_runs;
        dg[i].tile_begin = (uint32_t)tiles;
        dg[i].fchunk_begin = 0;
        tiles += (d.n_bases + TILE_BASES - 1) / TILE_BASES;
        max_bases = std::max<uint64_t>(max_bases, d.n_bases);
        if (tiles >> 31) return fail(ctx, MSBT_EINVAL, "batch too large (%llu tiles)", (unsigned long long)tiles);
    }

    if (!ctx->spec_is_default) {
        // a re-pinned spec (msbt_hash_spec): the generic kernel, whatever variant was asked for
        if (min_abund >= 2) return fail(ctx, MSBT_EINVAL, "min_abund >= 2 needs the default msbt_hash_spec in this version");
        size_t table_bytes = 0;
        int rc = stage_genome_table(ctx, dg, &table_bytes, stream);
        if (rc) return rc;
        for (uint32_t i = 0; i < n_genomes; i++)
            CU_TRY(ctx, cudaMemsetAsync(d_filters + dg[i].filter_word_off, 0, words * 8, stream));
        if (tiles == 0) return MSBT_OK;  // no genome holds a k-mer: the zeroed filters are the result
        const uint32_t key_len = ctx->spec.key_bytes == MSBT_KEY_BYTES_WHOLE_WORDS ? (k <= 32 ? 8u : 16u) : (k + 3u) / 4u;
        const int grid = (int)std::min<uint64_t>(tiles, (uint64_t)ctx->num_sms * 8);
        if (k <= 32)
            kmer_bloom_generic_kernel<1><<<grid, TILE_THREADS, 0, stream>>>(d_packed, d_run_ends, (const DevGenome *)ctx->d_scratch, n_genomes,
                                                                          (uint32_t)tiles, spec, key_len, ctx->spec.digest_half,
                                                                          ctx->spec.canonical, (uint32_t *)d_filters);
        else
            kmer_bloom_generic_kernel<2><<<grid, TILE_THREADS, 0, stream>>>(d_packed, d_run_ends, (const DevGenome *)ctx->d_scratch, n_genomes,
                                                                          (uint32_t)tiles, spec, key_len, ctx->spec.digest_half,
                                                                          ctx->spec.canonical, (uint32_t *)d_filters);
        LAUNCH_CHECK(ctx);
        return MSBT_OK;
    }

    // variant choice: the tiled path needs distinct output filters and u32 positions
    bool distinct = true;
    {
        std::vector<uint64_t> offs(n_genomes);
        for (uint32_t i = 0; i < n_genomes; i++) offs[i] = dg[i].filter_word_off;
        std::sort(offs.begin(), offs.end());
        for (uint32_t i = 1; i < n_genomes; i++) distinct = distinct && offs[i] != offs[i - 1];
    }
    const bool tiled_ok = distinct && min_abund <= 1 && num_bits <= (1ull << 32);
    if (variant >= 2 && !tiled_ok)
        return fail(ctx, MSBT_EINVAL, "variant %d needs distinct filter_index values, min_abund <= 1 and num_bits <= 2^32", variant);
    const bool fused_ok = tiled_ok && num_bits >= (1ull << 17);
    if (variant >= 3 && !fused_ok) return fail(ctx, MSBT_EINVAL, "variant %d needs num_bits >= 2^17", variant);
    if (fused_ok && variant == 4) {
        const int rc = build_phased(ctx, d_packed, d_run_ends, dg, tiles, spec, words, d_filters, stream);
        if (rc <= 0) return rc;
        if (tiles) return fail(ctx, MSBT_EINVAL, "variant 4: geometry not applicable to this batch");
    }
    // auto: the fused kernel from 2^25 bits up (34 us/genome at 5 Mbp for every size up to 2^28), the direct kernel below
    // (a filter of <= 2 MiB is L2 resident: 26-29 us/genome).  The tiled variant is never chosen automatically: with few
    // tiles it has no parallelism left (3.5 ms/genome at 2^22 bits).  tools/bench_small_filters.py
    if (fused_ok && (variant == 3 || (variant == 0 && num_bits >= (1ull << 25)))) {
        const int rc = build_fused(ctx, d_packed, d_run_ends, dg, tiles, spec, words, d_filters, stream);
        if (rc <= 0) return rc;
        if (variant == 3 && tiles) return fail(ctx, MSBT_EINVAL, "variant 3: geometry not applicable to this batch");
    }
    const bool tiled = tiled_ok && variant == 2;
    if (tiled) {
        const int rc = build_tiled(ctx, d_packed, d_run_ends, dg, tiles, spec, words, d_filters, stream);
        if (rc <= 0) return rc;
        if (variant == 2) return fail(ctx, MSBT_EINVAL, "variant 2: bucket scratch exceeds 2^32 entries per filter");
    }

    // Direct kernel in L2-sized groups: with every filter of a large batch zeroed up front, none of them is L2-resident by
    // the time its genome is hashed, and each RED becomes a DRAM sector read-modify-write (146 us/genome at 2^28 bits).
    // Zeroing a group of genomes whose filters fit L2 together right before hashing that group keeps the REDs in L2.
    uint64_t group_mb = 0;
    if (const char *e = getenv("MSBT_DIRECT_GROUP_MB")) group_mb = strtoull(e, nullptr, 10);
    if (group_mb && min_abund <= 1 && distinct && tiles && words * 8 <= (group_mb << 20) && (uint64_t)n_genomes * words * 8 > (group_mb << 20)) {
        size_t table_b = 0;
        int rc = stage_genome_table(ctx, dg, &table_b, stream);
        if (rc) return rc;
        const uint32_t per_group = (uint32_t)std::max<uint64_t>(1, (group_mb << 20) / (words * 8));
        for (uint32_t g0 = 0; g0 < n_genomes; g0 += per_group) {
            const uint32_t g1 = std::min(n_genomes, g0 + per_group);
            for (uint32_t i = g0; i < g1; i++) CU_TRY(ctx, cudaMemsetAsync(d_filters + dg[i].filter_word_off, 0, words * 8, stream));
            const uint64_t t0 = dg[g0].tile_begin, t1 = g1 < n_genomes ? dg[g1].tile_begin : tiles;
            rc = launch_direct(ctx, d_packed, d_run_ends, (const DevGenome *)ctx->d_scratch, n_genomes, t1, spec, d_filters, nullptr, stream, t0);
            if (rc) return rc;
        }
        return MSBT_OK;
    }

    // MSBT_PRE_ZERO_INLINE=1 (tuning): the `--min >= 2` build zeroes each filter on its genome's side stream right before the
    // emit pass instead of all of them up front
    const bool pre_path = min_abund >= 2 && k <= 31 && max_bases < (1ull << 30) && !getenv("MSBT_MIN_TABLE");
    const char *zi_env = getenv("MSBT_PRE_ZERO_INLINE");
    const bool pre_zero_inline = pre_path && distinct && zi_env && zi_env[0] == '1';
    // zero-fill every output filter (the reference always writes a complete new file)
    if (!pre_zero_inline) {
        // one memset per maximal contiguous span of output filters (ascending filter_index, dense stride)
        std::vector<uint64_t> offs(n_genomes);
        for (uint32_t i = 0; i < n_genomes; i++) offs[i] = dg[i].filter_word_off;
        std::sort(offs.begin(), offs.end());
        for (uint32_t i = 0; i < n_genomes;) {
            uint32_t j = i + 1;
            while (j < n_genomes && filter_stride_words == words && offs[j] == offs[j - 1] + words) j++;
            CU_TRY(ctx, cudaMemsetAsync(d_filters + offs[i], 0, (uint64_t)(j - i) * words * 8, stream));
            i = j;
        }
    }

    if (pre_path) {
        // prefiltered exact counting (see kmer_pre_mark_kernel); MSBT_MIN_TABLE=1 selects the plain count tables below
        uint64_t bits_per_base = 32;  // MSBT_PRE_BITS_PER_BASE (tuning): a smaller mark table leaves more of L2 to the other genomes in flight
        if (const char *e = getenv("MSBT_PRE_BITS_PER_BASE")) bits_per_base = std::min<uint64_t>(64, std::max<uint64_t>(1, strtoull(e, nullptr, 10)));
        uint64_t bit_count = 1ull << 16;
        while (bit_count < bits_per_base * max_bases && bit_count < (1ull << 33)) bit_count <<= 1;
        uint64_t max_slots = 1024;
        while (max_slots < 2 * max_bases) max_slots <<= 1;
        const size_t st_off = 0, bits_off = 256;
        const size_t cand_off = align_up(bits_off + bit_count / 8, 256);
        const size_t tab_off = align_up(cand_off + max_bases * 8, 256);
        const size_t cnt_off = align_up(tab_off + max_slots * 8, 256);
        const size_t cb_off = align_up(cnt_off + max_slots * 4, 256);
        const size_t slot_bytes = align_up(cb_off + PRE_CB_WORDS_MAX * 4, 256);
        // MSBT_PRE_STREAMS (1..4, default 3): genomes in flight at once, each with its own scratch slot and side stream
        int n_streams = 3;
        if (const char *e = getenv("MSBT_PRE_STREAMS")) n_streams = std::min(4, std::max(1, atoi(e)));
        n_streams = (int)std::min<uint32_t>((uint32_t)n_streams, n_genomes);
        int rc = ensure_scratch(ctx, slot_bytes * n_streams);
        if (rc) return rc;
        if (n_streams > 1) {
            if (!ctx->fork_ev) CU_TRY(ctx, cudaEventCreateWithFlags(&ctx->fork_ev, cudaEventDisableTiming));
            while (ctx->n_side < n_streams) {
                CU_TRY(ctx, cudaStreamCreateWithFlags(&ctx->side[ctx->n_side], cudaStreamNonBlocking));
                CU_TRY(ctx, cudaEventCreateWithFlags(&ctx->side_done[ctx->n_side], cudaEventDisableTiming));
                ctx->n_side++;
            }
            CU_TRY(ctx, cudaEventRecord(ctx->fork_ev, stream));  // the zeroed filters and the staged tables are ready
            for (int q = 0; q < n_streams; q++) CU_TRY(ctx, cudaStreamWaitEvent(ctx->side[q], ctx->fork_ev, 0));
        }
        // MSBT_PRE_COUNT_CB=0: the count pass without the shared-memory candidate bitmap (every k-mer reads the table);
        // tuning: MSBT_PRE_CB_LOG2 (16..20; <= 19 lets two CTAs share an SM), MSBT_PRE_CB_UNIT (1, 2, 4), MSBT_PRE_CB_MULT
        const char *cb_env = getenv("MSBT_PRE_COUNT_CB");
        const bool use_cb = !(cb_env && cb_env[0] == '0');
        uint32_t cb_log2 = 19, cb_ug = 1, cb_mult = 1;
        if (const char *e = getenv("MSBT_PRE_CB_LOG2")) cb_log2 = std::min<uint32_t>(PRE_CB_LOG2_MAX, std::max<uint32_t>(16, (uint32_t)atoi(e)));
        if (const char *e = getenv("MSBT_PRE_CB_UNIT")) cb_ug = atoi(e) == 4 ? 4 : (atoi(e) == 2 ? 2 : 1);
        if (const char *e = getenv("MSBT_PRE_CB_MULT")) cb_mult = e[0] == '1';
        const uint32_t cb_words = 1u << (cb_log2 - 5), cb_mask = (1u << cb_log2) - 1u;
        static bool cb_attr_set = false;
        if (use_cb && !cb_attr_set) {
            CU_TRY(ctx, cudaFuncSetAttribute(kmer_pre_count_cb_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(PRE_CB_WORDS_MAX * 4)));
            cb_attr_set = true;
        }
        uint32_t turn = 0;
        for (uint32_t i = 0; i < n_genomes; i++) {
            if (dg[i].n_bases < k) {
                if (pre_zero_inline) CU_TRY(ctx, cudaMemsetAsync(d_filters + dg[i].filter_word_off, 0, words * 8, stream));
                continue;
            }
            const int q = (int)(turn++ % (uint32_t)n_streams);
            cudaStream_t qs = n_streams > 1 ? ctx->side[q] : stream;
            char *S = (char *)ctx->d_scratch + (size_t)q * slot_bytes;
            PreState *st = (PreState *)(S + st_off);
            uint32_t *bits = (uint32_t *)(S + bits_off);
            uint64_t *cand = (uint64_t *)(S + cand_off);
            unsigned long long *tab = (unsigned long long *)(S + tab_off);
            uint32_t *cnt = (uint32_t *)(S + cnt_off);
            uint32_t *cbits = (uint32_t *)(S + cb_off);
            uint64_t gbits = 1ull << 16;
            while (gbits < bits_per_base * dg[i].n_bases && gbits < bit_count) gbits <<= 1;
            uint64_t gslots = 1024;
            while (gslots < 2ull * dg[i].n_bases) gslots <<= 1;
            CU_TRY(ctx, cudaMemsetAsync(S, 0, bits_off + gbits / 8, qs));  // state + bit table
            const uint32_t n_words = (dg[i].n_bases + 31) / 32;
            const int grid = (int)std::min<uint64_t>((n_words + TILE_THREADS - 1) / TILE_THREADS, (uint64_t)ctx->num_sms * 8);
            const int grid2 = ctx->num_sms * 8;
            kmer_pre_mark_kernel<<<grid, TILE_THREADS, 0, qs>>>(d_packed, d_run_ends, dg[i], k, bits, gbits - 1, cand, st);
            LAUNCH_CHECK(ctx);
            kmer_pre_prepare_kernel<<<grid2, 256, 0, qs>>>(st, tab, cnt, (uint32_t)gslots, cbits, cb_words);
            LAUNCH_CHECK(ctx);
            kmer_pre_insert_kernel<<<grid2, 256, 0, qs>>>(st, cand, tab, cbits, cb_mask, cb_mult);
            LAUNCH_CHECK(ctx);
            if (use_cb) {
                const uint64_t units = (uint64_t)n_words * (4 / cb_ug);
                const int grid3 = (int)std::min<uint64_t>((units + PRE_COUNT_THREADS - 1) / PRE_COUNT_THREADS,
                                                          (uint64_t)ctx->num_sms * (cb_log2 <= 19 ? 2 : 1));
                kmer_pre_count_cb_kernel<<<grid3, PRE_COUNT_THREADS, cb_words * 4, qs>>>(d_packed, d_run_ends, dg[i], k, st, tab, cnt, cbits,
                                                                                       cb_mask, cb_mult, cb_ug);
            } else {
                kmer_pre_count_kernel<<<grid, TILE_THREADS, 0, qs>>>(d_packed, d_run_ends, dg[i], k, st, tab, cnt);
            }
            LAUNCH_CHECK(ctx);
            if (pre_zero_inline) {
                if (zi_env[1] == 's' && (words & 1) == 0) {  // "1s": streaming stores
                    bf_zero_stream_kernel<<<ctx->num_sms * 2, 256, 0, qs>>>((uint4 *)(d_filters + dg[i].filter_word_off), words / 2);
                    LAUNCH_CHECK(ctx);
                } else CU_TRY(ctx, cudaMemsetAsync(d_filters + dg[i].filter_word_off, 0, words * 8, qs));
            }
            kmer_pre_emit_kernel<<<grid2, 256, 0, qs>>>(st, tab, cnt, min_abund, spec, (uint32_t *)(d_filters + dg[i].filter_word_off));
            LAUNCH_CHECK(ctx);
        }
        if (n_streams > 1) {
            for (int q = 0; q < n_streams; q++) {
                CU_TRY(ctx, cudaEventRecord(ctx->side_done[q], ctx->side[q]));
                CU_TRY(ctx, cudaStreamWaitEvent(stream, ctx->side_done[q], 0));
            }
        }
        return MSBT_OK;
    }
    if (min_abund >= 2 && min_abund <= 3 && k <= 31) {
        // compact table: 8 bytes per slot, load factor <= 2/3
        uint64_t max_slots = 1024;
        while (2 * max_slots < 3 * max_bases) max_slots <<= 1;
        int rc = ensure_scratch(ctx, max_slots * 8);
        if (rc) return rc;
        unsigned long long *tab = (unsigned long long *)ctx->d_scratch;
        for (uint32_t i = 0; i < n_genomes; i++) {
            if (dg[i].n_bases < k) continue;
            uint64_t slots = 1024;
            while (2 * slots < 3ull * dg[i].n_bases) slots <<= 1;
            CU_TRY(ctx, cudaMemsetAsync(tab, 0, slots * 8, stream));
            const uint32_t n_words = (dg[i].n_bases + 31) / 32;
            const int grid = (int)std::min<uint64_t>((n_words + TILE_THREADS - 1) / TILE_THREADS, (uint64_t)ctx->num_sms * 8);
            kmer_count2_insert_kernel<<<grid, TILE_THREADS, 0, stream>>>(d_packed, d_run_ends, dg[i], k, tab, slots - 1);
            LAUNCH_CHECK(ctx);
            const int grid2 = (int)std::min<uint64_t>((slots + 255) / 256, (uint64_t)ctx->num_sms * 8);
            kmer_count2_emit_kernel<<<grid2, 256, 0, stream>>>(tab, slots, min_abund, spec,
                                                               (uint32_t *)(d_filters + dg[i].filter_word_off));
            LAUNCH_CHECK(ctx);
        }
        return MSBT_OK;
    }
    if (min_abund >= 2) {
        uint64_t n_slots = 1024;
        while (n_slots < 2 * max_bases) n_slots <<= 1;
        size_t keys_b = n_slots * 8, cnt_b = n_slots * 4;
        int rc = ensure_scratch(ctx, keys_b + cnt_b);
        if (rc) return rc;
        unsigned long long *keys = (unsigned long long *)ctx->d_scratch;
        uint32_t *counts = (uint32_t *)((char *)ctx->d_scratch + keys_b);
        for (uint32_t i = 0; i < n_genomes; i++) {
            if (dg[i].n_bases < k) continue;
            uint64_t slots = 1024;
            while (slots < 2ull * dg[i].n_bases) slots <<= 1;
            CU_TRY(ctx, cudaMemsetAsync(keys, 0xFF, slots * 8, stream));
            CU_TRY(ctx, cudaMemsetAsync(counts, 0, slots * 4, stream));
            uint32_t n_words = (dg[i].n_bases + 31) / 32;
            int grid = (int)std::min<uint64_t>((n_words + TILE_THREADS - 1) / TILE_THREADS, (uint64_t)ctx->num_sms * 8);
            kmer_count_insert_kernel<<<grid, TILE_THREADS, 0, stream>>>(d_packed, d_run_ends, dg[i], k, keys, counts, slots - 1);
            LAUNCH_CHECK(ctx);
            int grid2 = (int)std::min<uint64_t>((slots + 255) / 256, (uint64_t)ctx->num_sms * 8);
            kmer_count_emit_kernel<<<grid2, 256, 0, stream>>>(keys, counts, slots, min_abund, spec,
                                                              (uint32_t *)(d_filters + dg[i].filter_word_off));
            LAUNCH_CHECK(ctx);
        }
        return MSBT_OK;
    }

    if (tiles == 0) return MSBT_OK;
    size_t table_b = align_up(sizeof(DevGenome) * n_genomes, 256);
    int rc = ensure_scratch(ctx, table_b);
    if (rc) return rc;
    rc = stage_begin(ctx, table_b);
    if (rc) return rc;
    rc = stage_to_device(ctx, ctx->d_scratch, dg.data(), sizeof(DevGenome) * n_genomes, 0, stream);
    if (rc) return rc;
    rc = stage_end(ctx, stream);
    if (rc) return rc;

    rc = launch_direct(ctx, d_packed, d_run_ends, (const DevGenome *)ctx->d_scratch, n_genomes, tiles, spec, d_filters, nullptr, stream);
    if (rc) return rc;
    return MSBT_OK;
}

// --------------------------------------------------------------------------- device FASTA ingest (msbt_ingest.cuh)
#include "msbt_ingest.cuh"

extern "C" int msbt_fasta_pack_device(msbt_ctx *ctx, const char *d_text, const uint64_t *h_offsets, uint32_t n_files,
                                      uint32_t min_run, uint64_t *d_packed, uint64_t packed_cap_words, uint32_t *d_run_ends,
                                      uint64_t run_cap, msbt_genome_desc *h_descs, void *stream_) {
    if (!ctx) return MSBT_EINVAL;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n_files == 0) return MSBT_OK;
    if (!h_offsets || !h_descs || !d_packed || !d_run_ends) return fail(ctx, MSBT_EINVAL, "fasta_pack_device: null argument");
    if (min_run > 64) return fail(ctx, MSBT_EINVAL, "fasta_pack_device: min_run > 64");
    CTX_DEVICE(ctx);
    std::vector<FpFile> files(n_files);
    uint64_t chunks = 0, words = 0, runs = 0;
    const uint64_t need = std::max<uint32_t>(min_run, 1);
    for (uint32_t i = 0; i < n_files; i++) {
        if (h_offsets[i + 1] < h_offsets[i]) return fail(ctx, MSBT_EINVAL, "fasta_pack_device: offsets must be non-decreasing");
        const uint64_t len = h_offsets[i + 1] - h_offsets[i];
        if (len >> 32) return fail(ctx, MSBT_EINVAL, "fasta_pack_device: file %u has >= 2^32 bytes", i);
        files[i].text_off = h_offsets[i];
        files[i].text_len = len;
        files[i].chunk_begin = (uint32_t)chunks;
        files[i].n_chunks = (uint32_t)((len + FP_CHUNK - 1) / FP_CHUNK);
        files[i].packed_word_off = words;
        files[i].run_off = runs;
        files[i].bt_off = 0;
        chunks += files[i].n_chunks;
        words += (len + 31) / 32 + MSBT_PACK_PAD_WORDS;  // upper bound: every byte a base
        runs += len / need + 1;
        if (chunks >> 31) return fail(ctx, MSBT_EINVAL, "fasta_pack_device: batch too large");
    }
    if (words > packed_cap_words) return fail(ctx, MSBT_ENOSPC, "fasta_pack_device: packed buffer needs %llu words", (unsigned long long)words);
    if (runs > run_cap) return fail(ctx, MSBT_ENOSPC, "fasta_pack_device: run table needs %llu entries", (unsigned long long)runs);
    if (chunks && !d_text) return fail(ctx, MSBT_EINVAL, "fasta_pack_device: null text");
    const uint32_t n_chunks = (uint32_t)chunks;
    // scratch: [files][summary][state][chunk_bases][chunk_breaks][totals 2 x u64][out_totals 2 x u64]
    const size_t f_b = align_up(sizeof(FpFile) * n_files, 256), c_b = align_up((size_t)n_chunks * 4 + 4, 256), t_b = align_up((size_t)n_files * 16, 256);
    int rc = ensure_scratch(ctx, f_b + 4 * c_b + 2 * t_b);
    if (rc) return rc;
    char *S = (char *)ctx->d_scratch;
    FpFile *d_files = (FpFile *)S;
    uint32_t *d_sum = (uint32_t *)(S + f_b), *d_state = (uint32_t *)(S + f_b + c_b);
    uint32_t *d_cb = (uint32_t *)(S + f_b + 2 * c_b), *d_ck = (uint32_t *)(S + f_b + 3 * c_b);
    unsigned long long *d_tot = (unsigned long long *)(S + f_b + 4 * c_b), *d_out = (unsigned long long *)(S + f_b + 4 * c_b + t_b);
    CU_TRY(ctx, cudaMemcpyAsync(d_files, files.data(), sizeof(FpFile) * n_files, cudaMemcpyHostToDevice, stream));
    CU_TRY(ctx, cudaMemsetAsync(d_packed, 0, words * 8, stream));
    const uint8_t *text = (const uint8_t *)d_text;
    std::vector<unsigned long long> tot(2 * (size_t)n_files, 0), out(2 * (size_t)n_files, 0);
    uint32_t *d_bt = nullptr;
    struct BtGuard {  // the run tables are freed on every exit path
        uint32_t *&p;
        ~BtGuard() { if (p) cudaFree(p); }
    } bt_guard{d_bt};
    if (n_chunks) {
        fp_line_summary_kernel<<<n_chunks, FP_THREADS, 0, stream>>>(text, d_files, n_files, n_chunks, d_sum);
        LAUNCH_CHECK(ctx);
        fp_line_state_kernel<<<n_files, 1024, 0, stream>>>(d_files, d_sum, d_state);
        LAUNCH_CHECK(ctx);
        fp_count_kernel<<<n_chunks, FP_THREADS, 0, stream>>>(text, d_files, n_files, n_chunks, d_state, d_cb, d_ck);
        LAUNCH_CHECK(ctx);
        fp_chunk_scan_kernel<<<n_files, 1024, 0, stream>>>(d_files, d_cb, d_ck, d_tot);
        LAUNCH_CHECK(ctx);
        CU_TRY(ctx, cudaMemcpyAsync(tot.data(), d_tot, sizeof(unsigned long long) * 2 * n_files, cudaMemcpyDeviceToHost, stream));
        CU_TRY(ctx, cudaStreamSynchronize(stream));
        uint64_t breaks = 0;
        for (uint32_t i = 0; i < n_files; i++) { files[i].bt_off = breaks; breaks += tot[2 * i + 1]; }
        // break table (breaks entries) and kp table (breaks + n_files entries)
        if (cudaMalloc(&d_bt, (2 * breaks + n_files + 1) * 4) != cudaSuccess) {
            cudaGetLastError();
            return fail(ctx, MSBT_ENOMEM, "fasta_pack_device: cudaMalloc of the run tables failed");
        }
        uint32_t *d_kp = d_bt + breaks;
        CU_TRY(ctx, cudaMemcpyAsync(d_files, files.data(), sizeof(FpFile) * n_files, cudaMemcpyHostToDevice, stream));
        fp_break_table_kernel<<<n_chunks, FP_THREADS, 0, stream>>>(text, d_files, n_files, n_chunks, d_state, d_cb, d_ck, d_bt);
        LAUNCH_CHECK(ctx);
        fp_run_scan_kernel<<<n_files, 1024, 0, stream>>>(d_files, d_tot, d_bt, min_run, d_kp, d_run_ends, d_out);
        LAUNCH_CHECK(ctx);
        fp_emit_kernel<<<n_chunks, FP_THREADS, 0, stream>>>(text, d_files, n_files, n_chunks, d_state, d_cb, d_ck, d_tot, d_bt, d_kp,
                                                          min_run, (unsigned long long *)d_packed);
        LAUNCH_CHECK(ctx);
        CU_TRY(ctx, cudaMemcpyAsync(out.data(), d_out, sizeof(unsigned long long) * 2 * n_files, cudaMemcpyDeviceToHost, stream));
        CU_TRY(ctx, cudaStreamSynchronize(stream));
    }
    for (uint32_t i = 0; i < n_files; i++) {
        h_descs[i].packed_word_off = files[i].packed_word_off;
        h_descs[i].n_bases = files[i].n_chunks ? out[2 * i] : 0;
        h_descs[i].run_off = files[i].run_off;
        h_descs[i].n_runs = files[i].n_chunks ? (uint32_t)out[2 * i + 1] : 0;
        h_descs[i].filter_index = i;
    }
    return MSBT_OK;
}

// --------------------------------------------------------------------------- K2 / K3 filter algebra

__device__ __forceinline__ uint4 ld_stream(const uint4 *p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

__device__ __forceinline__ uint32_t popc4(uint4 v) { return __popc(v.x) + __popc(v.y) + __popc(v.z) + __popc(v.w); }

__device__ __forceinline__ uint64_t block_sum_u64(uint64_t v) {
    __shared__ uint64_t warp_part[32];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) warp_part[wid] = v;
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    v = (threadIdx.x < nw) ? warp_part[threadIdx.x] : 0;
    if (wid == 0)
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;  // valid in thread 0
}

// dst = OR of n sources; `vec` = number of uint4 (two u64 words) per filter.
__global__ void __launch_bounds__(256)
bf_or_reduce_kernel(const uint4 *const *__restrict__ srcs, uint32_t n, uint64_t vec, uint4 *__restrict__ dst) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < vec; i += (uint64_t)gridDim.x * blockDim.x) {
        uint4 acc = make_uint4(0, 0, 0, 0);
        uint32_t s = 0;
        for (; s + 4 <= n; s += 4) {
            uint4 a = ld_stream(srcs[s] + i), b = ld_stream(srcs[s + 1] + i);
            uint4 c = ld_stream(srcs[s + 2] + i), d = ld_stream(srcs[s + 3] + i);
            acc.x |= a.x | b.x | c.x | d.x;
            acc.y |= a.y | b.y | c.y | d.y;
            acc.z |= a.z | b.z | c.z | d.z;
            acc.w |= a.w | b.w | c.w | d.w;
        }
        for (; s < n; s++) {
            uint4 a = ld_stream(srcs[s] + i);
            acc.x |= a.x; acc.y |= a.y; acc.z |= a.z; acc.w |= a.w;
        }
        dst[i] = acc;
    }
}

// out[0] += bits set in at least one of the n filters, out[1] += bits set in at least two: one streaming pass, nothing
// filter-sized is written (kopt.py: the number of distinct k-mers seen in >= 2 genomes).  `vec` = uint4 per filter.
__global__ void __launch_bounds__(256)
bf_occurrence_kernel(const uint4 *const *__restrict__ srcs, uint32_t n, uint64_t vec, uint64_t words,
                     unsigned long long *__restrict__ out) {
    uint64_t c1 = 0, c2 = 0;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < vec; i += (uint64_t)gridDim.x * blockDim.x) {
        uint4 once = make_uint4(0, 0, 0, 0), twice = make_uint4(0, 0, 0, 0);
        uint32_t s = 0;
        for (; s + 2 <= n; s += 2) {
            const uint4 a = ld_stream(srcs[s] + i), b = ld_stream(srcs[s + 1] + i);
            twice.x |= (once.x & (a.x | b.x)) | (a.x & b.x); once.x |= a.x | b.x;
            twice.y |= (once.y & (a.y | b.y)) | (a.y & b.y); once.y |= a.y | b.y;
            twice.z |= (once.z & (a.z | b.z)) | (a.z & b.z); once.z |= a.z | b.z;
            twice.w |= (once.w & (a.w | b.w)) | (a.w & b.w); once.w |= a.w | b.w;
        }
        if (s < n) {
            const uint4 a = ld_stream(srcs[s] + i);
            twice.x |= once.x & a.x; once.x |= a.x;
            twice.y |= once.y & a.y; once.y |= a.y;
            twice.z |= once.z & a.z; once.z |= a.z;
            twice.w |= once.w & a.w; once.w |= a.w;
        }
        c1 += popc4(once);
        c2 += popc4(twice);
    }
    if ((words & 1) && blockIdx.x == 0 && threadIdx.x == 0) {  // odd number of u64 words
        uint64_t once = 0, twice = 0;
        for (uint32_t s = 0; s < n; s++) {
            const uint64_t a = ((const uint64_t *)srcs[s])[words - 1];
            twice |= once & a;
            once |= a;
        }
        c1 += __popcll(once);
        c2 += __popcll(twice);
    }
    c1 = block_sum_u64(c1);
    c2 = block_sum_u64(c2);
    if (threadIdx.x == 0) {
        if (c1) atomicAdd(out + 0, (unsigned long long)c1);
        if (c2) atomicAdd(out + 1, (unsigned long long)c2);
    }
}

// scalar tail for an odd number of u64 words
__global__ void bf_or_tail_kernel(const uint64_t *const *__restrict__ srcs, uint32_t n, uint64_t w, uint64_t *__restrict__ dst) {
    uint64_t acc = 0;
    for (uint32_t s = 0; s < n; s++) acc |= srcs[s][w];
    dst[w] = acc;
}

// out[f] += popcount of a slice of filter f.  grid = (chunks, n)
__global__ void __launch_bounds__(256)
bf_popcount_kernel(const uint64_t *const *__restrict__ filters, uint64_t words, unsigned long long *__restrict__ out) {
    const uint64_t *__restrict__ f = filters[blockIdx.y];
    const uint64_t vec = words >> 1;
    const uint4 *fv = (const uint4 *)f;
    uint64_t c = 0;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < vec; i += (uint64_t)gridDim.x * blockDim.x)
        c += popc4(ld_stream(fv + i));
    if ((words & 1) && blockIdx.x == 0 && threadIdx.x == 0) c += __popcll(f[words - 1]);
    c = block_sum_u64(c);
    if (threadIdx.x == 0 && c) atomicAdd(out + blockIdx.y, (unsigned long long)c);
}

// One focus against TB targets per CTA pass: the focus slice is read once per TB targets.
constexpr int DIST_TB = 4;
__global__ void __launch_bounds__(256)
bf_dist_1vN_kernel(const uint64_t *__restrict__ focus, const uint64_t *const *__restrict__ targets, uint32_t n,
                   uint64_t words, unsigned long long *__restrict__ out_and, unsigned long long *__restrict__ out_or) {
    const uint32_t t0 = blockIdx.y * DIST_TB;
    const uint64_t vec = words >> 1;
    const uint4 *fv = (const uint4 *)focus;
    const uint4 *tv[DIST_TB];
#pragma unroll
    for (int t = 0; t < DIST_TB; t++) tv[t] = (const uint4 *)targets[min(t0 + t, n - 1)];
    uint64_t ca[DIST_TB] = {0}, co[DIST_TB] = {0};
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < vec; i += (uint64_t)gridDim.x * blockDim.x) {
        uint4 a = ld_stream(fv + i);
        uint4 b[DIST_TB];
#pragma unroll
        for (int t = 0; t < DIST_TB; t++) b[t] = ld_stream(tv[t] + i);
#pragma unroll
        for (int t = 0; t < DIST_TB; t++) {
            ca[t] += __popc(a.x & b[t].x) + __popc(a.y & b[t].y) + __popc(a.z & b[t].z) + __popc(a.w & b[t].w);
            co[t] += __popc(a.x | b[t].x) + __popc(a.y | b[t].y) + __popc(a.z | b[t].z) + __popc(a.w | b[t].w);
        }
    }
    if ((words & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
        uint64_t a = focus[words - 1];
#pragma unroll
        for (int t = 0; t < DIST_TB; t++) {
            uint64_t b = ((const uint64_t *)tv[t])[words - 1];
            ca[t] += __popcll(a & b);
            co[t] += __popcll(a | b);
        }
    }
#pragma unroll
    for (int t = 0; t < DIST_TB; t++) {
        uint64_t sa = block_sum_u64(ca[t]);
        uint64_t so = block_sum_u64(co[t]);
        if (threadIdx.x == 0 && t0 + t < n) {
            if (sa) atomicAdd(out_and + t0 + t, (unsigned long long)sa);
            if (so) atomicAdd(out_or + t0 + t, (unsigned long long)so);
        }
    }
}

// Arbitrary pair list: pair p = (a[p], b[p]); grid = (chunks, pairs).  The batched `dist` requests of profile_many (every MAG
// against its own handful of cluster representatives / centroids) -- one launch for all of them.
__global__ void __launch_bounds__(256)
bf_dist_pairs_kernel(const uint64_t *const *__restrict__ as, const uint64_t *const *__restrict__ bs, uint64_t words,
                     unsigned long long *__restrict__ out_and, unsigned long long *__restrict__ out_or) {
    const uint64_t *__restrict__ a = as[blockIdx.y], *__restrict__ b = bs[blockIdx.y];
    const uint64_t vec = words >> 1;
    const uint4 *av = (const uint4 *)a, *bv = (const uint4 *)b;
    uint64_t ca = 0, co = 0;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < vec; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint4 x = ld_stream(av + i), y = ld_stream(bv + i);
        ca += __popc(x.x & y.x) + __popc(x.y & y.y) + __popc(x.z & y.z) + __popc(x.w & y.w);
        co += __popc(x.x | y.x) + __popc(x.y | y.y) + __popc(x.z | y.z) + __popc(x.w | y.w);
    }
    if ((words & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
        ca += __popcll(a[words - 1] & b[words - 1]);
        co += __popcll(a[words - 1] | b[words - 1]);
    }
    ca = block_sum_u64(ca);
    co = block_sum_u64(co);
    if (threadIdx.x == 0) {
        if (ca) atomicAdd(out_and + blockIdx.y, (unsigned long long)ca);
        if (co) atomicAdd(out_or + blockIdx.y, (unsigned long long)co);
    }
}

// All pairs: CTA (bx, by, bz) handles filters I = by*AP_T.., J = bz*AP_T.. over word chunks bx.
// Each thread keeps AP_T x AP_T partial counts; the 2*AP_T slices are read once per tile pair.
constexpr int AP_T = 4;
__global__ void __launch_bounds__(256)
bf_allpairs_kernel(const uint64_t *const *__restrict__ filters, uint32_t n, uint64_t words,
                   unsigned long long *__restrict__ out) {
    const uint32_t i0 = blockIdx.y * AP_T, j0 = blockIdx.z * AP_T;
    if (j0 + AP_T <= i0) return;  // strictly-lower tiles: covered by symmetry
    const uint64_t vec = words >> 1;
    const uint4 *iv[AP_T], *jv[AP_T];
#pragma unroll
    for (int t = 0; t < AP_T; t++) {
        iv[t] = (const uint4 *)filters[min(i0 + t, n - 1)];
        jv[t] = (const uint4 *)filters[min(j0 + t, n - 1)];
    }
    uint32_t c[AP_T][AP_T] = {{0}};
    uint64_t acc[AP_T][AP_T] = {{0}};
    uint32_t iters = 0;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < vec; i += (uint64_t)gridDim.x * blockDim.x) {
        uint4 a[AP_T], b[AP_T];
#pragma unroll
        for (int t = 0; t < AP_T; t++) { a[t] = ld_stream(iv[t] + i); b[t] = ld_stream(jv[t] + i); }
#pragma unroll
        for (int x = 0; x < AP_T; x++)
#pragma unroll
            for (int y = 0; y < AP_T; y++)
                c[x][y] += __popc(a[x].x & b[y].x) + __popc(a[x].y & b[y].y) + __popc(a[x].z & b[y].z) + __popc(a[x].w & b[y].w);
        if (++iters == (1u << 24)) {  // 128 bits per iteration: flush long before u32 overflow
#pragma unroll
            for (int x = 0; x < AP_T; x++)
#pragma unroll
                for (int y = 0; y < AP_T; y++) { acc[x][y] += c[x][y]; c[x][y] = 0; }
            iters = 0;
        }
    }
    if ((words & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
#pragma unroll
        for (int x = 0; x < AP_T; x++)
#pragma unroll
            for (int y = 0; y < AP_T; y++)
                acc[x][y] += __popcll(((const uint64_t *)iv[x])[words - 1] & ((const uint64_t *)jv[y])[words - 1]);
    }
#pragma unroll
    for (int x = 0; x < AP_T; x++)
#pragma unroll
        for (int y = 0; y < AP_T; y++) {
            uint64_t s = block_sum_u64(acc[x][y] + c[x][y]);
            const uint32_t fi = i0 + x, fj = j0 + y;
            if (threadIdx.x == 0 && fi < n && fj < n && fi <= fj && s) {
                atomicAdd(out + (uint64_t)fi * n + fj, (unsigned long long)s);
                if (fi != fj) atomicAdd(out + (uint64_t)fj * n + fi, (unsigned long long)s);
            }
        }
}

// All pairs, carry-save version (the default).  POPC runs at 16 lanes/clk/SM, and the kernel above spends one POPC per
// 32 pair-bits.  Here a CTA stages one 1 KiB run of up to 2 x 16 filters in shared memory (every filter of the two
// blocks is read once per block pair) and a LANE owns a 2 x 2 BLOCK OF PAIRS (filters 2di, 2di+1 against 2dj, 2dj+1):
// per 4 words it reads four 16-byte operands from shared memory, ANDs them and feeds four carry-save accumulators
// (see apc_step for the depth of the tree).  The 2 x 2
// blocking halves the shared-memory bytes per pair-word (a warp-wide 16-byte load takes four data cycles whatever it
// broadcasts), which is what bounds a one-pair-per-lane version (measured: 4.9 ms for 64 filters of 2^28 bits against
// 5.3 ms for the POPC kernel).  When a block pair has fewer than 32 lane blocks (a 10-genome species has 15), the spare
// lanes take other word ranges of the tile, so the warp stays full.
constexpr int APC_B = 16;                    // filters per block (8 duos)
constexpr int APC_TW = 256;                  // u32 words of every filter per tile (1 KiB)
constexpr int APC_THREADS = 256;
constexpr int APC_WARPS = APC_THREADS / 32;
constexpr size_t APC_SMEM_BYTES = (size_t)2 * 2 * APC_B * (APC_TW + 4) * 4;  // two tile buffers; rows are 16-byte aligned and 4 banks apart

__device__ __forceinline__ void csa32(uint32_t &hi, uint32_t &lo, uint32_t a, uint32_t b, uint32_t c) {
    uint32_t s_, c_;  // the outputs may alias the inputs
    asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(s_) : "r"(a), "r"(b), "r"(c));  // a ^ b ^ c
    asm("lop3.b32 %0, %1, %2, %3, 0xE8;" : "=r"(c_) : "r"(a), "r"(b), "r"(c));  // majority
    lo = s_;
    hi = c_;
}

struct ApcPair {
    uint32_t ones, twos_count;
};

// 4 words of one pair.  The tree is deliberately only one level deep: LOP3/AND issue at 64 lanes/clk/SM (the ALU pipe)
// and POPC at 16, so a word costs max(0.5 * ALU ops, 2 * POPCs) cycles per warp -- 2.0 with a POPC per word, 1.44 with
// a full 16-word Harley-Seal tree (2.9 ALU ops per word; measured: ALU pipe 86 % busy, 4.6 ms for 64 filters of 2^28
// bits), and 1.0 on BOTH pipes with one carry-save step per two words (2 ALU ops and half a POPC per word).
__device__ __forceinline__ void apc_step(ApcPair &s, const uint4 &a, const uint4 &b) {
    uint32_t t;
    csa32(t, s.ones, s.ones, a.x & b.x, a.y & b.y);
    s.twos_count += __popc(t);
    csa32(t, s.ones, s.ones, a.z & b.z, a.w & b.w);
    s.twos_count += __popc(t);
}

// Row of filter f of a block of nf filters inside the tile: first members of the duos in rows 0 .., second members
// behind them.  Rows are 4 banks apart, so the eight lanes of a quarter warp, which read the same member of eight
// different duos, hit 8 x 4 distinct banks.
__device__ __forceinline__ uint32_t apc_row(uint32_t f, uint32_t nf) { return (f & 1) * ((nf + 1) / 2) + (f >> 1); }

__device__ __forceinline__ void apc_step4(ApcPair (&st)[4], const uint32_t *a0, const uint32_t *a1, const uint32_t *b0, const uint32_t *b1) {
    const uint4 va0 = *(const uint4 *)a0, va1 = *(const uint4 *)a1, vb0 = *(const uint4 *)b0, vb1 = *(const uint4 *)b1;
    apc_step(st[0], va0, vb0);
    apc_step(st[1], va0, vb1);
    apc_step(st[2], va1, vb0);
    apc_step(st[3], va1, vb1);
}

__global__ void __launch_bounds__(APC_THREADS, 3)
bf_allpairs_csa_kernel(const uint64_t *const *__restrict__ filters, uint32_t n, uint64_t words, uint32_t pair_base,
                       unsigned long long *__restrict__ out, const uint64_t *const *__restrict__ filters_b = nullptr,
                       uint32_t n_b = 0) {
    // filters_b != nullptr: RECTANGULAR mode -- every filter of `filters` (n of them) against every filter of `filters_b`
    // (n_b), out is n x n_b; block pairs are (I over A, J over B), none of them diagonal.
    extern __shared__ __align__(16) uint32_t apc_tile[];            // rows of tw + 4 words: I block, then J block
    __shared__ const uint64_t *sm_ptr[2 * APC_B];
    __shared__ uint32_t sm_dst[2 * APC_B];
    __shared__ uint8_t sm_lb[64];                                   // lane block -> (di << 4) | dj
    __shared__ unsigned long long sm_cnt[64 * 4];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // block pair (I <= J) from the linear index
    const bool rect = filters_b != nullptr;
    const uint32_t nb = (n + APC_B - 1) / APC_B;
    uint32_t I = 0, t = pair_base + blockIdx.y;  // (more than 65,535 block pairs take several launches)
    uint32_t J;
    if (rect) {
        const uint32_t nbb = (n_b + APC_B - 1) / APC_B;
        I = t / nbb;
        J = t - I * nbb;
    } else {
        while (t >= nb - I) { t -= nb - I; I++; }
        J = I + t;
    }
    const bool diag = !rect && I == J;
    const uint32_t ni = min((uint32_t)APC_B, n - I * APC_B), nj = min((uint32_t)APC_B, (rect ? n_b : n) - J * APC_B);
    const uint32_t di_n = (ni + 1) / 2, dj_n = (nj + 1) / 2;
    const uint32_t LB = diag ? di_n * (di_n + 1) / 2 : di_n * dj_n;  // lane blocks: 1 .. 64
    // words of every filter per tile: the tile buffer holds 32 rows of 1 KiB; with fewer filters the rows get longer
    // (10 filters: 2 KiB each), so that a CTA keeps the same number of bytes in flight per load phase
    const uint32_t nslices = diag ? ni : ni + nj;
    const uint32_t tw = nslices <= 8 ? 4 * APC_TW : nslices <= 16 ? 2 * APC_TW : APC_TW, stride = tw + 4, tv = tw / 4;
    const uint32_t tsh = 31 - __clz(tv);  // tv is a power of two
    // lanes: LB >= 32 -> R rounds of 32 lane blocks; LB < 32 -> one round, `sub` word ranges side by side in a warp
    const uint32_t R = (LB + 31) / 32;
    uint32_t sub = 1;
    while (2 * sub * LB <= 32 && 2 * sub * 16 <= tw) sub <<= 1;
    // work items (one per warp): R rounds x S spans of the tile; a lane's word range = span / sub words, >= 16
    uint32_t S = 1;
    while (R * S * 2 <= APC_WARPS && S * sub * 2 * 16 <= tw) S <<= 1;
    const uint32_t lane_words = tw / (S * sub);
    if (threadIdx.x < 2 * APC_B) {
        const uint32_t sl = threadIdx.x;
        const uint64_t *p = nullptr;
        if (sl < ni) p = filters[I * APC_B + sl];
        else if (!diag && sl < ni + nj) p = (rect ? filters_b : filters)[J * APC_B + sl - ni];
        sm_ptr[sl] = p;
        // byte offset of the slice's row: the I block's filters first (duo members interleaved by apc_row), then the J block's
        sm_dst[sl] = (sl < ni ? apc_row(sl, ni) : ni + apc_row(min(sl - ni, nj - 1), nj)) * stride * 4;
    }
    if (threadIdx.x < 64) {
        uint32_t code = 0;
        const uint32_t p = threadIdx.x;
        if (p < LB) {
            if (diag) {  // row-major upper triangle of duos, with the diagonal
                uint32_t i = 0, r = p;
                while (r >= di_n - i) { r -= di_n - i; i++; }
                code = (i << 4) | (i + r);
            } else code = ((p / dj_n) << 4) | (p % dj_n);
        }
        sm_lb[p] = (uint8_t)code;
    }
    sm_cnt[threadIdx.x] = 0;  // 256 threads, 256 counters
    __syncthreads();
    // this lane's 2 x 2 block and word range
    const bool warp_act = warp < R * S;
    const uint32_t r = warp_act ? warp % R : 0, sp = warp_act ? warp / R : 0;
    const uint32_t lb = sub > 1 ? lane % LB : r * 32 + lane, g = sub > 1 ? lane / LB : 0;
    const bool lane_act = warp_act && lb < LB && g < sub;
    const uint32_t code = lane_act ? sm_lb[lb] : 0;
    const uint32_t di = code >> 4, dj = code & 15;
    const uint32_t woff = lane_act ? (sp * sub + g) * lane_words : 0;
    // rows: the I block's filters first (duo members interleaved by apc_row), then the J block's
    const uint32_t jbase = diag ? 0 : ni;
    const uint32_t *a0 = apc_tile + apc_row(min(2 * di, ni - 1), ni) * stride + woff;
    const uint32_t *a1 = apc_tile + apc_row(min(2 * di + 1, ni - 1), ni) * stride + woff;
    const uint32_t *b0 = apc_tile + (jbase + apc_row(min(2 * dj, nj - 1), nj)) * stride + woff;
    const uint32_t *b1 = apc_tile + (jbase + apc_row(min(2 * dj + 1, nj - 1), nj)) * stride + woff;
    ApcPair st[4];
#pragma unroll
    for (int k = 0; k < 4; k++) st[k].ones = st[k].twos_count = 0;
    const uint64_t vec_full = words >> 1;                           // whole uint4 of every filter
    const uint64_t n_vec = (words + 1) >> 1;
    const uint64_t n_tiles = (n_vec + tv - 1) / tv;
    constexpr int LD_PASSES = 2 * APC_B * (APC_TW / 4) / APC_THREADS;  // 8
    constexpr uint32_t BUF_WORDS = 2 * APC_B * (APC_TW + 4);
    // two tile buffers: the 16-byte cp.async copies of the next tile are in flight while this one is consumed
    // Thread t copies 16-byte chunk c = t mod tv of slices s0, s0 + ds, ... (tv divides the CTA size, so c is the same
    // in every pass): per pass one pointer and one row offset come from shared memory, and the end-of-filter test
    // is one test per tile.
    const uint32_t c = threadIdx.x & (tv - 1), s0 = threadIdx.x >> tsh, ds = APC_THREADS >> tsh;
    const uint32_t tile_sa = (uint32_t)__cvta_generic_to_shared(apc_tile) + 16 * c;
    auto issue_tile = [&](uint64_t tile, uint32_t buf) {
        const uint64_t gv = tile * tv + c;
        if (gv < vec_full) {
            const uint32_t sa = tile_sa + buf * (BUF_WORDS * 4);
#pragma unroll
            for (int k = 0; k < LD_PASSES; k++) {
                const uint32_t sl = s0 + k * ds;
                if (sl < nslices)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa + sm_dst[sl]), "l"((const uint4 *)sm_ptr[sl] + gv) : "memory");
            }
        } else {  // past the last whole vector: the odd last word, or zeros
#pragma unroll 1
            for (int k = 0; k < LD_PASSES; k++) {
                const uint32_t sl = s0 + k * ds;
                if (sl >= nslices) break;
                uint4 z = make_uint4(0, 0, 0, 0);
                if (gv < n_vec) { const uint64_t w = sm_ptr[sl][words - 1]; z.x = (uint32_t)w; z.y = (uint32_t)(w >> 32); }
                *(uint4 *)((char *)apc_tile + buf * (BUF_WORDS * 4) + sm_dst[sl] + 16 * c) = z;
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    uint32_t buf = 0;
    if (blockIdx.x < n_tiles) issue_tile(blockIdx.x, 0);
    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const uint64_t next = tile + gridDim.x;
        if (next < n_tiles) {
            issue_tile(next, buf ^ 1);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();  // this tile has landed for every thread
        if (warp_act) {
            const uint32_t bo = buf * BUF_WORDS;
#pragma unroll 4
            for (uint32_t w = 0; w < lane_words; w += 4) apc_step4(st, a0 + bo + w, a1 + bo + w, b0 + bo + w, b1 + bo + w);
        }
        __syncthreads();  // consumed: the next iteration refills this buffer
        buf ^= 1;
    }
    // fold the carry-save planes; the word ranges of a pair meet in shared memory; one global atomic per pair and CTA
    if (lane_act) {
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const unsigned long long c = 2ull * st[k].twos_count + __popc(st[k].ones);
            if (c) atomicAdd(&sm_cnt[lb * 4 + k], c);
        }
    }
    __syncthreads();
    {
        const uint32_t p = threadIdx.x >> 2, k = threadIdx.x & 3;   // 64 lane blocks x 4 pairs = 256 threads
        const unsigned long long c = sm_cnt[threadIdx.x];
        if (p < LB && c) {
            const uint32_t cd = sm_lb[p];
            const uint32_t li = 2 * (cd >> 4) + (k >> 1), lj = 2 * (cd & 15) + (k & 1);
            const uint32_t fi = I * APC_B + li, fj = J * APC_B + lj;
            if (rect) {
                if (li < ni && lj < nj) atomicAdd(out + (uint64_t)fi * n_b + fj, c);
            } else if (li < ni && lj < nj && fi <= fj) {
                atomicAdd(out + (uint64_t)fi * n + fj, c);
                if (fi != fj) atomicAdd(out + (uint64_t)fj * n + fi, c);
            }
        }
    }
}

// Upload a host array of n device pointers into scratch; returns the device copy.
static int upload_ptrs(msbt_ctx *ctx, const uint64_t *const *h_ptrs, uint32_t n, cudaStream_t stream,
                       const uint64_t *const **d_out) {
    size_t bytes = align_up((size_t)n * sizeof(void *), 256);
    int rc = ensure_scratch(ctx, bytes);
    if (rc) return rc;
    rc = stage_begin(ctx, bytes);
    if (rc) return rc;
    rc = stage_to_device(ctx, ctx->d_scratch, h_ptrs, (size_t)n * sizeof(void *), 0, stream);
    if (rc) return rc;
    rc = stage_end(ctx, stream);
    if (rc) return rc;
    *d_out = (const uint64_t *const *)ctx->d_scratch;
    return MSBT_OK;
}

static int check_ptrs(msbt_ctx *ctx, const uint64_t *const *h, uint32_t n) {
    if (!h) return fail(ctx, MSBT_EINVAL, "null pointer table");
    for (uint32_t i = 0; i < n; i++) {
        if (!h[i]) return fail(ctx, MSBT_EINVAL, "filter %u is a null pointer", i);
        if (((uintptr_t)h[i]) & 15) return fail(ctx, MSBT_EINVAL, "filter %u is not 16-byte aligned", i);
    }
    return MSBT_OK;
}

static int stream_grid(const msbt_ctx *ctx, uint64_t vec, int per_sm) {
    uint64_t want = (vec + 255) / 256;
    return (int)std::max<uint64_t>(1, std::min<uint64_t>(want, (uint64_t)ctx->num_sms * per_sm));
}

extern "C" int msbt_bf_or_reduce(msbt_ctx *ctx, const uint64_t *const *h_srcs, uint32_t n, uint64_t words,
                                 uint64_t *d_dst, void *stream_) {
    if (!ctx) return MSBT_EINVAL;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n == 0 || words == 0 || !d_dst) return fail(ctx, MSBT_EINVAL, "bf_or_reduce: empty input");
    int rc = check_ptrs(ctx, h_srcs, n);
    if (rc) return rc;
    if (((uintptr_t)d_dst) & 15) return fail(ctx, MSBT_EINVAL, "destination is not 16-byte aligned");
    CTX_DEVICE(ctx);
    const uint64_t *const *d_srcs;
    rc = upload_ptrs(ctx, h_srcs, n, stream, &d_srcs);
    if (rc) return rc;
    const uint64_t vec = words >> 1;
    if (vec) {
        bf_or_reduce_kernel<<<stream_grid(ctx, vec, 8), 256, 0, stream>>>((const uint4 *const *)d_srcs, n, vec, (uint4 *)d_dst);
        LAUNCH_CHECK(ctx);
    }
    if (words & 1) {
        bf_or_tail_kernel<<<1, 1, 0, stream>>>(d_srcs, n, words - 1, d_dst);
        LAUNCH_CHECK(ctx);
    }
    return MSBT_OK;
}

extern "C" int msbt_bf_occurrence(msbt_ctx *ctx, const uint64_t *const *h_filters, uint32_t n, uint64_t words,
                                  uint64_t *d_out, void *stream_) {
    if (!ctx) return MSBT_EINVAL;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (words == 0 || !d_out) return fail(ctx, MSBT_EINVAL, "bf_occurrence: empty input");
    CTX_DEVICE(ctx);
    CU_TRY(ctx, cudaMemsetAsync(d_out, 0, 16, stream));
    if (n == 0) return MSBT_OK;
    int rc = check_ptrs(ctx, h_filters, n);
    if (rc) return rc;
    const uint64_t *const *d_f;
    rc = upload_ptrs(ctx, h_filters, n, stream, &d_f);
    if (rc) return rc;
    const uint64_t vec = words >> 1;
    bf_occurrence_kernel<<<stream_grid(ctx, std::max<uint64_t>(vec, 1), 8), 256, 0, stream>>>((const uint4 *const *)d_f, n, vec, words,
                                                                                             (unsigned long long *)d_out);
    LAUNCH_CHECK(ctx);
    return MSBT_OK;
}

extern "C" int msbt_bf_popcount(msbt_ctx *ctx, const uint64_t *const *h_filters, uint32_t n, uint64_t words,
                                uint64_t *d_out, void *stream_) {
    if (!ctx) return MSBT_EINVAL;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n == 0) return MSBT_OK;
    if (words == 0 || !d_out) return fail(ctx, MSBT_EINVAL, "bf_popcount: empty input");
    int rc = check_ptrs(ctx, h_filters, n);
    if (rc) return rc;
    CTX_DEVICE(ctx);
    const uint64_t *const *d_f;
    rc = upload_ptrs(ctx, h_filters, n, stream, &d_f);
    if (rc) return rc;
    CU_TRY(ctx, cudaMemsetAsync(d_out, 0, (size_t)n * 8, stream));
    const uint64_t vec = std::max<uint64_t>(words >> 1, 1);
    uint64_t gx = std::max<uint64_t>(1, std::min<uint64_t>((vec + 255) / 256, std::max<uint64_t>(1, ((uint64_t)ctx->num_sms * 32 + n - 1) / n)));
    for (uint32_t off = 0; off < n; off += 65535) {
        uint32_t cnt = std::min<uint32_t>(65535, n - off);
        bf_popcount_kernel<<<dim3((unsigned)gx, cnt), 256, 0, stream>>>(d_f + off, words, (unsigned long long *)d_out + off);
        LAUNCH_CHECK(ctx);
    }
    return MSBT_OK;
}

extern "C" int msbt_bf_and_popcount_1vN(msbt_ctx *ctx, const uint64_t *d_focus, const uint64_t *const *h_targets,
                                        uint32_t n, uint64_t words, uint64_t *d_intersect, uint64_t *d_union,
                                        void *stream_) {
    if (!ctx) return MSBT_EINVAL;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n == 0) return MSBT_OK;
    if (words == 0 || !d_focus || !d_intersect || !d_union) return fail(ctx, MSBT_EINVAL, "bf_and_popcount_1vN: null/empty input");
    if (((uintptr_t)d_focus) & 15) return fail(ctx, MSBT_EINVAL, "focus is not 16-byte aligned");
    int rc = check_ptrs(ctx, h_targets, n);
    if (rc) return rc;
    CTX_DEVICE(ctx);
    const uint64_t *const *d_t;
    rc = upload_ptrs(ctx, h_targets, n, stream, &d_t);
    if (rc) return rc;
    CU_TRY(ctx, cudaMemsetAsync(d_intersect, 0, (size_t)n * 8, stream));
    CU_TRY(ctx, cudaMemsetAsync(d_union, 0, (size_t)n * 8, stream));
    const uint64_t vec = std::max<uint64_t>(words >> 1, 1);
    const uint32_t groups = (n + DIST_TB - 1) / DIST_TB;
    uint64_t gx = std::max<uint64_t>(1, std::min<uint64_t>((vec + 255) / 256, std::max<uint64_t>(1, (uint64_t)ctx->num_sms * 8 / groups)));
    for (uint32_t off = 0; off < groups; off += 65535) {
        uint32_t cnt = std::min<uint32_t>(65535, groups - off);
        uint32_t first = off * DIST_TB;
        bf_dist_1vN_kernel<<<dim3((unsigned)gx, cnt), 256, 0, stream>>>(d_focus, d_t + first, n - first, words,
                                                                       (unsigned long long *)d_intersect + first,
                                                                       (unsigned long long *)d_union + first);
        LAUNCH_CHECK(ctx);
    }
    return MSBT_OK;
}

extern "C" int msbt_bf_and_popcount_pairs(msbt_ctx *ctx, const uint64_t *const *h_a, const uint64_t *const *h_b, uint32_t n_pairs,
                                          uint64_t words, uint64_t *d_intersect, uint64_t *d_union, void *stream_) {
    if (!ctx) return MSBT_EINVAL;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n_pairs == 0) return MSBT_OK;
    if (words == 0 || !d_intersect || !d_union) return fail(ctx, MSBT_EINVAL, "bf_and_popcount_pairs: null/empty input");
    int rc = check_ptrs(ctx, h_a, n_pairs);
    if (rc) return rc;
    rc = check_ptrs(ctx, h_b, n_pairs);
    if (rc) return rc;
    CTX_DEVICE(ctx);
    // both pointer tables side by side in scratch
    const size_t tab = align_up((size_t)n_pairs * sizeof(void *), 256);
    rc = ensure_scratch(ctx, 2 * tab);
    if (rc) return rc;
    rc = stage_begin(ctx, 2 * tab);
    if (rc) return rc;
    rc = stage_to_device(ctx, ctx->d_scratch, h_a, (size_t)n_pairs * sizeof(void *), 0, stream);
    if (rc) return rc;
    rc = stage_to_device(ctx, (char *)ctx->d_scratch + tab, h_b, (size_t)n_pairs * sizeof(void *), tab, stream);
    if (rc) return rc;
    rc = stage_end(ctx, stream);
    if (rc) return rc;
    const uint64_t *const *d_a = (const uint64_t *const *)ctx->d_scratch;
    const uint64_t *const *d_b = (const uint64_t *const *)((char *)ctx->d_scratch + tab);
    CU_TRY(ctx, cudaMemsetAsync(d_intersect, 0, (size_t)n_pairs * 8, stream));
    CU_TRY(ctx, cudaMemsetAsync(d_union, 0, (size_t)n_pairs * 8, stream));
    const uint64_t vec = std::max<uint64_t>(words >> 1, 1);
    const uint64_t gx = std::max<uint64_t>(1, std::min<uint64_t>((vec + 255) / 256, std::max<uint64_t>(1, ((uint64_t)ctx->num_sms * 16 + n_pairs - 1) / n_pairs)));
    for (uint32_t off = 0; off < n_pairs; off += 65535) {
        const uint32_t cnt = std::min<uint32_t>(65535, n_pairs - off);
        bf_dist_pairs_kernel<<<dim3((unsigned)gx, cnt), 256, 0, stream>>>(d_a + off, d_b + off, words,
                                                                         (unsigned long long *)d_intersect + off,
                                                                         (unsigned long long *)d_union + off);
        LAUNCH_CHECK(ctx);
    }
    return MSBT_OK;
}

extern "C" int msbt_bf_and_popcount_allpairs(msbt_ctx *ctx, const uint64_t *const *h_filters, uint32_t n,
                                             uint64_t words, uint64_t *d_intersect, void *stream_) {
    if (!ctx) return MSBT_EINVAL;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n == 0) return MSBT_OK;
    if (words == 0 || !d_intersect) return fail(ctx, MSBT_EINVAL, "bf_and_popcount_allpairs: null/empty input");
    int rc = check_ptrs(ctx, h_filters, n);
    if (rc) return rc;
    CTX_DEVICE(ctx);
    int variant = 0;  // 0 = carry-save kernel (default), 1 = the plain POPC kernel (kept as the cross-check)
    if (const char *e = getenv("MSBT_ALLPAIRS")) variant = atoi(e);
    const uint64_t *const *d_f;
    if (variant == 1) {
        const uint32_t tiles = (n + AP_T - 1) / AP_T;
        if (tiles > 65535) return fail(ctx, MSBT_EINVAL, "allpairs: n=%u too large for one call", n);
        rc = upload_ptrs(ctx, h_filters, n, stream, &d_f);
        if (rc) return rc;
        CU_TRY(ctx, cudaMemsetAsync(d_intersect, 0, (size_t)n * n * 8, stream));
        const uint64_t vec = std::max<uint64_t>(words >> 1, 1);
        const uint64_t pair_tiles = (uint64_t)tiles * (tiles + 1) / 2;
        uint64_t gx = std::max<uint64_t>(1, std::min<uint64_t>((vec + 255) / 256, std::max<uint64_t>(1, (uint64_t)ctx->num_sms * 8 / pair_tiles)));
        bf_allpairs_kernel<<<dim3((unsigned)gx, tiles, tiles), 256, 0, stream>>>(d_f, n, words, (unsigned long long *)d_intersect);
        LAUNCH_CHECK(ctx);
        return MSBT_OK;
    }
    const uint64_t nb = (n + APC_B - 1) / APC_B, block_pairs = nb * (nb + 1) / 2;
    if (block_pairs > 0xFFFFFFFFull) return fail(ctx, MSBT_EINVAL, "allpairs: n=%u too large for one call", n);
    rc = upload_ptrs(ctx, h_filters, n, stream, &d_f);
    if (rc) return rc;
    CU_TRY(ctx, cudaMemsetAsync(d_intersect, 0, (size_t)n * n * 8, stream));
    const uint64_t n_tiles = (((words + 1) >> 1) + APC_TW / 4 - 1) / (APC_TW / 4);  // upper bound (small n: longer tiles)
    const uint64_t gx = std::max<uint64_t>(1, std::min<uint64_t>(n_tiles, std::max<uint64_t>(1, (uint64_t)ctx->num_sms * 6 / block_pairs)));
    CU_TRY(ctx, cudaFuncSetAttribute(bf_allpairs_csa_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)APC_SMEM_BYTES));
    // gridDim.y is limited to 65,535: n > 5,776 filters (361 blocks of 16) take one launch per 65,535 block pairs
    for (uint64_t base = 0; base < block_pairs; base += 65535) {
        const unsigned cnt = (unsigned)std::min<uint64_t>(65535, block_pairs - base);
        bf_allpairs_csa_kernel<<<dim3((unsigned)gx, cnt), APC_THREADS, APC_SMEM_BYTES, stream>>>(
            d_f, n, words, (uint32_t)base, (unsigned long long *)d_intersect);
        LAUNCH_CHECK(ctx);
    }
    return MSBT_OK;
}

// Every filter of list A against every filter of list B: d_intersect is a dense n_a x n_b row-major u64 matrix.  The same
// carry-save kernel as the square case, block pairs (I over A) x (J over B).
extern "C" int msbt_bf_and_popcount_rect(msbt_ctx *ctx, const uint64_t *const *h_a, uint32_t n_a, const uint64_t *const *h_b,
                                         uint32_t n_b, uint64_t words, uint64_t *d_intersect, void *stream_) {
    if (!ctx) return MSBT_EINVAL;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n_a == 0 || n_b == 0) return MSBT_OK;
    if (words == 0 || !d_intersect) return fail(ctx, MSBT_EINVAL, "bf_and_popcount_rect: null/empty input");
    int rc = check_ptrs(ctx, h_a, n_a);
    if (rc) return rc;
    rc = check_ptrs(ctx, h_b, n_b);
    if (rc) return rc;
    CTX_DEVICE(ctx);
    const size_t tab_a = align_up((size_t)n_a * sizeof(void *), 256), tab_b = align_up((size_t)n_b * sizeof(void *), 256);
    rc = ensure_scratch(ctx, tab_a + tab_b);
    if (rc) return rc;
    rc = stage_begin(ctx, tab_a + tab_b);
    if (rc) return rc;
    rc = stage_to_device(ctx, ctx->d_scratch, h_a, (size_t)n_a * sizeof(void *), 0, stream);
    if (rc) return rc;
    rc = stage_to_device(ctx, (char *)ctx->d_scratch + tab_a, h_b, (size_t)n_b * sizeof(void *), tab_a, stream);
    if (rc) return rc;
    rc = stage_end(ctx, stream);
    if (rc) return rc;
    const uint64_t *const *d_a = (const uint64_t *const *)ctx->d_scratch;
    const uint64_t *const *d_b = (const uint64_t *const *)((char *)ctx->d_scratch + tab_a);
    CU_TRY(ctx, cudaMemsetAsync(d_intersect, 0, (size_t)n_a * n_b * 8, stream));
    const uint64_t nba = (n_a + APC_B - 1) / APC_B, nbb = (n_b + APC_B - 1) / APC_B, block_pairs = nba * nbb;
    if (block_pairs > 0xFFFFFFFFull) return fail(ctx, MSBT_EINVAL, "bf_and_popcount_rect: too many block pairs");
    const uint64_t n_tiles = (((words + 1) >> 1) + APC_TW / 4 - 1) / (APC_TW / 4);
    const uint64_t gx = std::max<uint64_t>(1, std::min<uint64_t>(n_tiles, std::max<uint64_t>(1, (uint64_t)ctx->num_sms * 6 / block_pairs)));
    CU_TRY(ctx, cudaFuncSetAttribute(bf_allpairs_csa_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)APC_SMEM_BYTES));
    for (uint64_t base = 0; base < block_pairs; base += 65535) {
        const unsigned cnt = (unsigned)std::min<uint64_t>(65535, block_pairs - base);
        bf_allpairs_csa_kernel<<<dim3((unsigned)gx, cnt), APC_THREADS, APC_SMEM_BYTES, stream>>>(
            d_a, n_a, words, (uint32_t)base, (unsigned long long *)d_intersect, d_b, n_b);
        LAUNCH_CHECK(ctx);
    }
    return MSBT_OK;
}

// --------------------------------------------------------------------------- host: sparse filter transfer
// Ascending set-bit positions (the output of msbt_bf_extract_positions*, copied device -> host) -> the dense filter words.
// At the densities the path works at (5 Mbp into 2^30 bits: 0.47 %) a filter's position list is 1/6 of its dense size,
// so it is the cheaper thing to move over PCIe; the `.bf` payload is rebuilt on the host, one 32 KiB block at a time
// (zeroed and filled while it is in L1).
extern "C" int msbt_positions_expand(const uint32_t *positions, uint64_t n, uint64_t num_bits, uint64_t *words) {
    if (!words || (n && !positions)) return MSBT_EINVAL;
    const uint64_t n_words = (num_bits + 63) / 64;
    constexpr uint64_t BLOCK = 4096;  // words: one 32 KiB block is assembled in an L1-resident scratch buffer ...
    alignas(64) uint64_t scratch[BLOCK];
    const bool stream_out = (((uintptr_t)words) & 15) == 0;  // ... and leaves with non-temporal stores (no read-for-ownership:
    uint64_t i = 0;                                           // the host memory traffic is the filter's size, written once)
    for (uint64_t w0 = 0; w0 < n_words; w0 += BLOCK) {
        const uint64_t w1 = std::min(n_words, w0 + BLOCK), nw = w1 - w0;
        memset(scratch, 0, nw * 8);
        const uint64_t bit_end = w1 * 64;
        while (i < n && positions[i] < bit_end) {
            const uint64_t p = positions[i++];
            if (p < w0 * 64) return MSBT_EINVAL;  // not ascending
            scratch[(p >> 6) - w0] |= 1ull << (p & 63);
        }
        uint64_t *dst = words + w0;
        uint64_t k = 0;
        if (stream_out) {
            for (; k + 2 <= nw; k += 2) _mm_stream_si128((__m128i *)(dst + k), _mm_load_si128((const __m128i *)(scratch + k)));
        }
        for (; k < nw; k++) dst[k] = scratch[k];
    }
    _mm_sfence();
    if (i != n) return MSBT_EINVAL;  // a position beyond the last word
    if (n && (num_bits & 63) && positions[n - 1] >= num_bits) return MSBT_EINVAL;
    return MSBT_OK;
}

// --------------------------------------------------------------------------- K4: query

constexpr int EX_THREADS = 256;
constexpr int EX_WORDS_PER_THREAD = 4;  // u64 words
constexpr int EX_CHUNK = EX_THREADS * EX_WORDS_PER_THREAD;

// pass 1: per-chunk popcounts
__global__ void __launch_bounds__(EX_THREADS)
bf_extract_count_kernel(const uint64_t *__restrict__ f, uint64_t words, unsigned long long *__restrict__ chunk_counts) {
    const uint64_t base = (uint64_t)blockIdx.x * EX_CHUNK + (uint64_t)threadIdx.x * EX_WORDS_PER_THREAD;
    uint64_t c = 0;
#pragma unroll
    for (int i = 0; i < EX_WORDS_PER_THREAD; i++)
        if (base + i < words) c += __popcll(f[base + i]);
    c = block_sum_u64(c);
    if (threadIdx.x == 0) chunk_counts[blockIdx.x] = c;
}

// pass 2: exclusive scan of chunk counts by one CTA (in place); total -> *d_count
__global__ void __launch_bounds__(1024)
bf_extract_scan_kernel(unsigned long long *__restrict__ chunk_counts, uint64_t n_chunks, unsigned long long *__restrict__ d_count) {
    __shared__ unsigned long long part[1024];
    const uint64_t per = (n_chunks + 1023) / 1024;
    const uint64_t b = threadIdx.x * per, e = min(b + per, n_chunks);
    unsigned long long s = 0;
    for (uint64_t i = b; i < e; i++) s += chunk_counts[i];
    part[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long run = 0;
        for (int i = 0; i < 1024; i++) { unsigned long long t = part[i]; part[i] = run; run += t; }
        *d_count = run;
    }
    __syncthreads();
    unsigned long long run = part[threadIdx.x];
    for (uint64_t i = b; i < e; i++) { unsigned long long t = chunk_counts[i]; chunk_counts[i] = run; run += t; }
}

// pass 3: ordered emission
__global__ void __launch_bounds__(EX_THREADS)
bf_extract_emit_kernel(const uint64_t *__restrict__ f, uint64_t words, const unsigned long long *__restrict__ chunk_offsets,
                       uint32_t *__restrict__ out, uint64_t cap) {
    __shared__ uint32_t warp_tot[EX_THREADS / 32];
    const uint64_t base = (uint64_t)blockIdx.x * EX_CHUNK + (uint64_t)threadIdx.x * EX_WORDS_PER_THREAD;
    uint64_t v[EX_WORDS_PER_THREAD];
    uint32_t mine = 0;
#pragma unroll
    for (int i = 0; i < EX_WORDS_PER_THREAD; i++) {
        v[i] = (base + i < words) ? f[base + i] : 0;
        mine += __popcll(v[i]);
    }
    // exclusive prefix inside the CTA
    uint32_t incl = mine;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) warp_tot[wid] = incl;
    __syncthreads();
    uint32_t woff = 0;
    for (int w2 = 0; w2 < wid; w2++) woff += warp_tot[w2];
    uint64_t o = chunk_offsets[blockIdx.x] + woff + (incl - mine);
#pragma unroll
    for (int i = 0; i < EX_WORDS_PER_THREAD; i++) {
        uint64_t x = v[i];
        while (x) {
            int b = __ffsll((long long)x) - 1;
            x &= x - 1;
            if (o < cap) out[o] = (uint32_t)(((base + i) << 6) + b);
            o++;
        }
    }
}

extern "C" int msbt_bf_extract_positions(msbt_ctx *ctx, const uint64_t *d_filter, uint64_t words,
                                         uint32_t *d_positions, uint64_t cap, uint64_t *d_count, void *stream_) {
    if (!ctx) return MSBT_EINVAL;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!d_filter || !d_count || words == 0) return fail(ctx, MSBT_EINVAL, "bf_extract_positions: null/empty input");
    if (words > (1ull << 26)) return fail(ctx, MSBT_EINVAL, "bf_extract_positions: positions are u32 (num_bits <= 2^32)");
    if (cap && !d_positions) return fail(ctx, MSBT_EINVAL, "bf_extract_positions: null output");
    CTX_DEVICE(ctx);
    const uint64_t n_chunks = (words + EX_CHUNK - 1) / EX_CHUNK;
    // chunk table lives past the first 64 KiB of scratch (pointer tables use the front)
    const size_t off = 1 << 16;
    int rc = ensure_scratch(ctx, off + n_chunks * 8);
    if (rc) return rc;
    unsigned long long *chunks = (unsigned long long *)((char *)ctx->d_scratch + off);
    bf_extract_count_kernel<<<(unsigned)n_chunks, EX_THREADS, 0, stream>>>(d_filter, words, chunks);
    LAUNCH_CHECK(ctx);
    bf_extract_scan_kernel<<<1, 1024, 0, stream>>>(chunks, n_chunks, (unsigned long long *)d_count);
    LAUNCH_CHECK(ctx);
    if (cap) {
        bf_extract_emit_kernel<<<(unsigned)n_chunks, EX_THREADS, 0, stream>>>(d_filter, words, chunks, d_positions, cap);
        LAUNCH_CHECK(ctx);
    }
    return MSBT_OK;
}

// ---- batched extraction: all query filters of a batch in three launches, no host round trip in between.
// grid (chunks, filters): per-chunk popcounts
__global__ void __launch_bounds__(EX_THREADS)
bf_extract_count_batch_kernel(const uint64_t *const *__restrict__ filters, uint64_t words, uint32_t n_chunks,
                              uint32_t *__restrict__ chunk_counts) {
    const uint64_t *__restrict__ f = filters[blockIdx.y];
    const uint64_t base = (uint64_t)blockIdx.x * EX_CHUNK + (uint64_t)threadIdx.x * EX_WORDS_PER_THREAD;
    uint64_t c = 0;
#pragma unroll
    for (int i = 0; i < EX_WORDS_PER_THREAD; i++)
        if (base + i < words) c += __popcll(f[base + i]);
    c = block_sum_u64(c);
    if (threadIdx.x == 0) chunk_counts[(uint64_t)blockIdx.y * n_chunks + blockIdx.x] = (uint32_t)c;
}

// one CTA per filter: exclusive scan of its chunk counts in place, total -> totals[f]
__global__ void __launch_bounds__(1024)
bf_extract_scan_batch_kernel(uint32_t *__restrict__ chunk_counts, uint32_t n_chunks, unsigned long long *__restrict__ totals) {
    __shared__ unsigned long long part[1024];
    uint32_t *cc = chunk_counts + (uint64_t)blockIdx.x * n_chunks;
    const uint32_t per = (n_chunks + 1023) / 1024;
    const uint32_t b = threadIdx.x * per, e = min(b + per, n_chunks);
    unsigned long long s = 0;
    for (uint32_t i = b; i < e; i++) s += cc[i];
    part[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long run = 0;
        for (int i = 0; i < 1024; i++) { const unsigned long long t = part[i]; part[i] = run; run += t; }
        totals[blockIdx.x] = run;
    }
    __syncthreads();
    unsigned long long run = part[threadIdx.x];
    for (uint32_t i = b; i < e; i++) { const uint32_t t = cc[i]; cc[i] = (uint32_t)run; run += t; }  // per-filter offsets < 2^32
}

// single CTA: offsets[f] = sum of totals[0..f), offsets[n] = grand total
__global__ void __launch_bounds__(1024)
bf_extract_offsets_kernel(const unsigned long long *__restrict__ totals, uint32_t n, unsigned long long *__restrict__ offsets) {
    __shared__ unsigned long long part[1024];
    const uint32_t per = (n + 1023) / 1024;
    const uint32_t b = threadIdx.x * per, e = min(b + per, n);
    unsigned long long s = 0;
    for (uint32_t i = b; i < e; i++) s += totals[i];
    part[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long run = 0;
        for (int i = 0; i < 1024; i++) { const unsigned long long t = part[i]; part[i] = run; run += t; }
        offsets[n] = run;
    }
    __syncthreads();
    unsigned long long run = part[threadIdx.x];
    for (uint32_t i = b; i < e; i++) { offsets[i] = run; run += totals[i]; }
}

// grid (chunks, filters): ordered emission into the concatenated output
__global__ void __launch_bounds__(EX_THREADS)
bf_extract_emit_batch_kernel(const uint64_t *const *__restrict__ filters, uint64_t words, uint32_t n_chunks,
                             const uint32_t *__restrict__ chunk_offsets, const unsigned long long *__restrict__ offsets,
                             uint32_t *__restrict__ out, uint64_t cap) {
    __shared__ uint32_t warp_tot[EX_THREADS / 32];
    const uint64_t *__restrict__ f = filters[blockIdx.y];
    const uint64_t base = (uint64_t)blockIdx.x * EX_CHUNK + (uint64_t)threadIdx.x * EX_WORDS_PER_THREAD;
    uint64_t v[EX_WORDS_PER_THREAD];
    uint32_t mine = 0;
#pragma unroll
    for (int i = 0; i < EX_WORDS_PER_THREAD; i++) {
        v[i] = (base + i < words) ? f[base + i] : 0;
        mine += __popcll(v[i]);
    }
    uint32_t incl = mine;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) warp_tot[wid] = incl;
    __syncthreads();
    uint32_t woff = 0;
    for (int w2 = 0; w2 < wid; w2++) woff += warp_tot[w2];
    uint64_t o = offsets[blockIdx.y] + chunk_offsets[(uint64_t)blockIdx.y * n_chunks + blockIdx.x] + woff + (incl - mine);
#pragma unroll
    for (int i = 0; i < EX_WORDS_PER_THREAD; i++) {
        uint64_t x = v[i];
        while (x) {
            const int bit = __ffsll((long long)x) - 1;
            x &= x - 1;
            if (o < cap) out[o] = (uint32_t)(((base + i) << 6) + bit);
            o++;
        }
    }
}

extern "C" int msbt_bf_extract_positions_batch(msbt_ctx *ctx, const uint64_t *const *h_filters, uint32_t n, uint64_t words,
                                               uint32_t *d_positions, uint64_t cap, uint64_t *d_offsets, void *stream_) {
    if (!ctx) return MSBT_EINVAL;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n == 0) return MSBT_OK;
    if (!d_offsets || words == 0) return fail(ctx, MSBT_EINVAL, "bf_extract_positions_batch: null/empty input");
    if (words > (1ull << 26)) return fail(ctx, MSBT_EINVAL, "bf_extract_positions_batch: positions are u32 (num_bits <= 2^32)");
    if (n > 65535) return fail(ctx, MSBT_EINVAL, "bf_extract_positions_batch: at most 65535 filters per call");
    if (cap && !d_positions) return fail(ctx, MSBT_EINVAL, "bf_extract_positions_batch: null output");
    int rc = check_ptrs(ctx, h_filters, n);
    if (rc) return rc;
    CTX_DEVICE(ctx);
    const uint32_t n_chunks = (uint32_t)((words + EX_CHUNK - 1) / EX_CHUNK);
    // scratch: [pointer table][totals n x u64][chunk counts n x n_chunks x u32]
    const size_t ptr_b = align_up((size_t)n * sizeof(void *), 256);
    const size_t tot_b = align_up((size_t)n * 8, 256);
    rc = ensure_scratch(ctx, ptr_b + tot_b + (size_t)n * n_chunks * 4);
    if (rc) return rc;
    const uint64_t *const *d_f;
    rc = upload_ptrs(ctx, h_filters, n, stream, &d_f);
    if (rc) return rc;
    unsigned long long *totals = (unsigned long long *)((char *)ctx->d_scratch + ptr_b);
    uint32_t *chunks = (uint32_t *)((char *)ctx->d_scratch + ptr_b + tot_b);
    bf_extract_count_batch_kernel<<<dim3(n_chunks, n), EX_THREADS, 0, stream>>>(d_f, words, n_chunks, chunks);
    LAUNCH_CHECK(ctx);
    bf_extract_scan_batch_kernel<<<n, 1024, 0, stream>>>(chunks, n_chunks, totals);
    LAUNCH_CHECK(ctx);
    bf_extract_offsets_kernel<<<1, 1024, 0, stream>>>(totals, n, (unsigned long long *)d_offsets);
    LAUNCH_CHECK(ctx);
    if (cap) {
        bf_extract_emit_batch_kernel<<<dim3(n_chunks, n), EX_THREADS, 0, stream>>>(d_f, words, n_chunks, chunks,
                                                                                  (const unsigned long long *)d_offsets, d_positions, cap);
        LAUNCH_CHECK(ctx);
    }
    return MSBT_OK;
}

// Row-major probe: CTA (leaf-group, query).  Each thread walks positions with stride and
// tests QL leaves per position so the position stream is read once per QL leaves.
constexpr int QL = 4;
__global__ void __launch_bounds__(256)
query_rowmajor_kernel(const uint64_t *const *__restrict__ leaves, uint32_t n_leaves, const uint32_t *__restrict__ positions,
                      const ulonglong2 *__restrict__ q_rng, uint32_t *__restrict__ hits) {
    const uint32_t q = blockIdx.y, l0 = blockIdx.x * QL;
    const uint64_t b = q_rng[q].x, e = q_rng[q].y;
    const uint32_t *lf[QL];
#pragma unroll
    for (int t = 0; t < QL; t++) lf[t] = (const uint32_t *)leaves[min(l0 + t, n_leaves - 1)];
    uint32_t c[QL] = {0};
    for (uint64_t i = b + threadIdx.x; i < e; i += blockDim.x) {
        const uint32_t p = positions[i];
#pragma unroll
        for (int t = 0; t < QL; t++) c[t] += (__ldg(lf[t] + (p >> 5)) >> (p & 31)) & 1u;
    }
#pragma unroll
    for (int t = 0; t < QL; t++) {
        uint64_t s = block_sum_u64(c[t]);
        if (threadIdx.x == 0 && l0 + t < n_leaves) hits[(uint64_t)q * n_leaves + l0 + t] = (uint32_t)s;
    }
}

// Query q owns positions [begin_q, end_q) of d_positions.  The (begin, end) pairs go to device scratch at `scratch_off`.
static int upload_ranges(msbt_ctx *ctx, const uint64_t *h_rng, uint32_t n_queries, size_t scratch_off,
                         cudaStream_t stream, const ulonglong2 **d_out) {
    size_t bytes = (size_t)n_queries * 16;
    int rc = ensure_scratch(ctx, scratch_off + bytes);
    if (rc) return rc;
    // staging area after the pointer table region
    rc = stage_to_device(ctx, (char *)ctx->d_scratch + scratch_off, h_rng, bytes, scratch_off, stream);
    if (rc) return rc;
    *d_out = (const ulonglong2 *)((char *)ctx->d_scratch + scratch_off);
    return MSBT_OK;
}

static int ranges_from_offsets(msbt_ctx *ctx, const uint64_t *h_off, uint32_t n_queries, std::vector<uint64_t> &rng) {
    if (!h_off) return fail(ctx, MSBT_EINVAL, "query: null offsets");
    rng.resize((size_t)n_queries * 2);
    for (uint32_t q = 0; q < n_queries; q++) {
        if (h_off[q + 1] < h_off[q]) return fail(ctx, MSBT_EINVAL, "query offsets must be non-decreasing");
        rng[2 * q] = h_off[q];
        rng[2 * q + 1] = h_off[q + 1];
    }
    return MSBT_OK;
}

extern "C" int msbt_query_batch_ranges(msbt_ctx *ctx, const uint64_t *const *h_leaves, uint32_t n_leaves,
                                       const uint32_t *d_positions, const uint64_t *h_query_ranges, uint32_t n_queries,
                                       uint32_t *d_hits, void *stream_) {
    if (!ctx) return MSBT_EINVAL;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n_queries == 0 || n_leaves == 0) return MSBT_OK;
    if (!h_query_ranges || !d_hits) return fail(ctx, MSBT_EINVAL, "query_batch: null argument");
    int rc = check_ptrs(ctx, h_leaves, n_leaves);
    if (rc) return rc;
    bool any = false;
    for (uint32_t q = 0; q < n_queries; q++) {
        if (h_query_ranges[2 * q + 1] < h_query_ranges[2 * q]) return fail(ctx, MSBT_EINVAL, "query %u: range end before begin", q);
        any = any || h_query_ranges[2 * q + 1] > h_query_ranges[2 * q];
    }
    if (any && !d_positions) return fail(ctx, MSBT_EINVAL, "query_batch: null positions");
    CTX_DEVICE(ctx);
    const size_t ptr_b = align_up((size_t)n_leaves * sizeof(void *), 256);
    const uint32_t groups = (n_leaves + QL - 1) / QL;
    // gridDim.y is limited to 65,535 queries per launch
    for (uint32_t q0 = 0; q0 < n_queries; q0 += 65535) {
        const uint32_t nq = std::min<uint32_t>(65535, n_queries - q0);
        const size_t rng_b = (size_t)nq * 16;
        rc = ensure_scratch(ctx, ptr_b + rng_b);
        if (rc) return rc;
        rc = stage_begin(ctx, ptr_b + rng_b);
        if (rc) return rc;
        rc = stage_to_device(ctx, ctx->d_scratch, h_leaves, (size_t)n_leaves * sizeof(void *), 0, stream);
        if (rc) return rc;
        const ulonglong2 *d_rng;
        rc = upload_ranges(ctx, h_query_ranges + 2 * (size_t)q0, nq, ptr_b, stream, &d_rng);
        if (rc) return rc;
        rc = stage_end(ctx, stream);
        if (rc) return rc;
        query_rowmajor_kernel<<<dim3(groups, nq), 256, 0, stream>>>((const uint64_t *const *)ctx->d_scratch, n_leaves,
                                                                  d_positions, d_rng, d_hits + (size_t)q0 * n_leaves);
        LAUNCH_CHECK(ctx);
    }
    return MSBT_OK;
}

extern "C" int msbt_query_batch(msbt_ctx *ctx, const uint64_t *const *h_leaves, uint32_t n_leaves,
                                const uint32_t *d_positions, const uint64_t *h_query_offsets, uint32_t n_queries,
                                uint32_t *d_hits, void *stream_) {
    if (!ctx) return MSBT_EINVAL;
    if (n_queries == 0 || n_leaves == 0) return MSBT_OK;
    std::vector<uint64_t> rng;
    int rc = ranges_from_offsets(ctx, h_query_offsets, n_queries, rng);
    if (rc) return rc;
    return msbt_query_batch_ranges(ctx, h_leaves, n_leaves, d_positions, rng.data(), n_queries, d_hits, stream_);
}

// Bit-slice transpose.  sliced[(p - row_begin) * ls + (l >> 5)] bit (l & 31) = bit p of leaf l.
// CTA tile = up to 512 leaves x 1024 rows: every leaf contributes one full 128-byte line per iteration (with one 32-byte
// sector per leaf the memory system fetched 3.2 x the leaf bytes from DRAM: profiles/r01_ncu_k4_summary.txt), the
// 32 x 32 bit blocks are transposed in registers with a 5-stage shuffle butterfly, and every warp streams its 32 rows
// out as 64-byte row segments through a private staging buffer -- read every leaf once, write the sliced matrix once.
// 100 KiB of shared memory per CTA: two CTAs per SM, one loading while the other one transposes.
constexpr int BS_WORDS = 32;                      // u32 words (= 1024 rows, 128 bytes) of every leaf per CTA iteration
constexpr int BS_COLS = 16;                       // column words (= 512 leaves) per CTA
constexpr int BS_LEAVES = BS_COLS * 32;
constexpr int BS_THREADS = 512;
constexpr int BS_WARPS = BS_THREADS / 32;
constexpr int BS_IN_STRIDE = BS_WORDS + 1;        // smem padding: conflict-free column reads
constexpr int BS_OUT_STRIDE = BS_COLS + 1;
constexpr size_t BS_SMEM_BYTES = ((size_t)BS_LEAVES * BS_IN_STRIDE + (size_t)BS_WARPS * 32 * BS_OUT_STRIDE) * 4;

// 32 x 32 bit transpose across a warp: lane r holds row r; returns column r (bit c = old row c, bit r).
// Stage J (mask M = the low J bits of every 2J-bit group): a lane of the lower half keeps (x & M) and takes
// (t & M) << J from its partner, a lane of the upper half keeps (x & ~M) and takes (t & ~M) >> J.  Both are
// (x & K) | (rot(t) & ~K) with a per-lane keep mask K and a per-lane rotation (left by J or by 32 - J; the bits a
// rotation wraps around fall outside ~K), so a stage is SHFL + SHF + LOP3 with no predication.
struct T32Lane {
    uint32_t keep[5], rot[5];
    __device__ __forceinline__ explicit T32Lane(uint32_t lane) {
        const uint32_t m[5] = {0x0000FFFFu, 0x00FF00FFu, 0x0F0F0F0Fu, 0x33333333u, 0x55555555u};
#pragma unroll
        for (int s = 0; s < 5; s++) {
            const uint32_t j = 16u >> s;
            const bool up = (lane & j) != 0;
            keep[s] = up ? ~m[s] : m[s];
            rot[s] = up ? 32u - j : j;
        }
    }
};

__device__ __forceinline__ uint32_t transpose32(uint32_t x, const T32Lane &tl) {
#pragma unroll
    for (int s = 0; s < 5; s++) {
        const uint32_t t = __shfl_xor_sync(0xffffffffu, x, 16 >> s);
        const uint32_t r = __funnelshift_l(t, t, tl.rot[s]);
        asm("lop3.b32 %0, %1, %2, %3, 0xCA;" : "=r"(x) : "r"(tl.keep[s]), "r"(x), "r"(r));  // keep ? x : r
    }
    return x;
}

// two independent transposes interleaved (the five stages of one are a dependent chain of shuffles)
__device__ __forceinline__ void transpose32x2(uint32_t &x, uint32_t &y, const T32Lane &tl) {
#pragma unroll
    for (int s = 0; s < 5; s++) {
        const uint32_t tx = __shfl_xor_sync(0xffffffffu, x, 16 >> s), ty = __shfl_xor_sync(0xffffffffu, y, 16 >> s);
        const uint32_t rx = __funnelshift_l(tx, tx, tl.rot[s]), ry = __funnelshift_l(ty, ty, tl.rot[s]);
        asm("lop3.b32 %0, %1, %2, %3, 0xCA;" : "=r"(x) : "r"(tl.keep[s]), "r"(x), "r"(rx));
        asm("lop3.b32 %0, %1, %2, %3, 0xCA;" : "=r"(y) : "r"(tl.keep[s]), "r"(y), "r"(ry));
    }
}

// One warp: transpose row word ww of the `cols` leaf groups held in `in`, stage the 32 rows x cols words in the warp's
// `out` buffer and write them as row segments (two rows per store instruction when cols == BS_COLS).
template <bool FULL>
__device__ __forceinline__ void bitslice_warp(const uint32_t *__restrict__ in, uint32_t *__restrict__ out, uint32_t ww, uint32_t cols,
                                              uint32_t lane, const T32Lane &tl, uint32_t *__restrict__ dst_row0, uint32_t ls,
                                              uint32_t rows_valid) {
    const uint32_t *src = in + lane * BS_IN_STRIDE + ww;
    if (FULL) {
#pragma unroll
        for (int lg = 0; lg < BS_COLS; lg += 2) {
            uint32_t x = src[lg * 32 * BS_IN_STRIDE], y = src[(lg + 1) * 32 * BS_IN_STRIDE];
            transpose32x2(x, y, tl);
            out[lane * BS_OUT_STRIDE + lg] = x;
            out[lane * BS_OUT_STRIDE + lg + 1] = y;
        }
        __syncwarp();
        // a store instruction writes rows i and i + 16 (16 * BS_OUT_STRIDE = 16 mod 32: the two half-warps read
        // disjoint banks of the staging buffer)
        constexpr int RPI = 32 / BS_COLS, RGAP = 32 / RPI;  // rows per instruction, distance between them
        const uint32_t c = lane % BS_COLS, r0 = (lane / BS_COLS) * RGAP;
        uint32_t *d = dst_row0 + (size_t)r0 * ls + c;
        const uint32_t *o = out + r0 * BS_OUT_STRIDE + c;
#pragma unroll
        for (int i = 0; i < RGAP; i++) {
            if (r0 + i < rows_valid) d[(size_t)i * ls] = o[i * BS_OUT_STRIDE];
        }
    } else {
        for (uint32_t lg = 0; lg < cols; lg++) out[lane * BS_OUT_STRIDE + lg] = transpose32(src[lg * 32 * BS_IN_STRIDE], tl);
        __syncwarp();
        for (uint32_t idx = lane; idx < 32 * cols; idx += 32) {
            const uint32_t r = idx / cols, c = idx - r * cols;
            if (r < rows_valid) dst_row0[(size_t)r * ls + c] = out[r * BS_OUT_STRIDE + c];
        }
    }
    __syncwarp();
}

__global__ void __launch_bounds__(BS_THREADS, 2)
bitslice_kernel(const uint64_t *const *__restrict__ leaves, uint32_t n_leaves, uint64_t row_begin, uint64_t row_end,
                uint32_t lw, uint32_t ls, uint32_t *__restrict__ sliced) {
    extern __shared__ uint32_t bs_smem[];
    uint32_t *in = bs_smem;                                     // [BS_LEAVES][BS_IN_STRIDE]
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t *out = bs_smem + BS_LEAVES * BS_IN_STRIDE + warp * 32 * BS_OUT_STRIDE;  // per warp: [32][BS_OUT_STRIDE]
    const uint32_t leaf0 = blockIdx.y * BS_LEAVES;
    const uint32_t lg0 = leaf0 >> 5;
    const uint32_t cols = min((uint32_t)BS_COLS, lw - lg0);     // column words of this tile
    const uint64_t n_rowwords = (row_end - row_begin + 31) >> 5;
    const uint64_t first_word = row_begin >> 5;                 // row_begin is a multiple of 32
    const bool vec_ok = (first_word & 3) == 0;                  // leaf pointers are 16-byte aligned (checked by the host)
    const T32Lane tl(lane);
    for (uint64_t wb = (uint64_t)blockIdx.x * BS_WORDS; wb < n_rowwords; wb += (uint64_t)gridDim.x * BS_WORDS) {
        __syncthreads();
        if (vec_ok && wb + BS_WORDS <= n_rowwords) {
            // load: 8 lanes per leaf, one 16-byte load each (a full 128-byte line per leaf), 64 leaves per pass;
            // all loads of a thread are issued before the first one is used
            constexpr int PASSES = BS_LEAVES / (BS_THREADS / 8);
            const uint32_t q = threadIdx.x & 7;
            uint4 v[PASSES];
#pragma unroll
            for (int p = 0; p < PASSES; p++) {
                const uint32_t l = (threadIdx.x >> 3) + p * (BS_THREADS / 8), leaf = leaf0 + l;
                v[p] = make_uint4(0, 0, 0, 0);
                if (l < cols * 32 && leaf < n_leaves) v[p] = ld_stream((const uint4 *)((const uint32_t *)leaves[leaf] + first_word + wb) + q);
            }
#pragma unroll
            for (int p = 0; p < PASSES; p++) {
                const uint32_t l = (threadIdx.x >> 3) + p * (BS_THREADS / 8);
                uint32_t *d = in + l * BS_IN_STRIDE + 4 * q;
                d[0] = v[p].x; d[1] = v[p].y; d[2] = v[p].z; d[3] = v[p].w;
            }
        } else {
            // general (unaligned first word or the last, partial tile): a warp reads the 32 words of one leaf
            for (uint32_t l = warp; l < cols * 32; l += BS_WARPS) {
                const uint32_t leaf = leaf0 + l;
                uint32_t x = 0;
                if (leaf < n_leaves && wb + lane < n_rowwords) x = __ldg((const uint32_t *)leaves[leaf] + first_word + wb + lane);
                in[l * BS_IN_STRIDE + lane] = x;
            }
        }
        __syncthreads();
        // transpose + store: warp w handles row words w, w + BS_WARPS of every leaf group
        for (uint32_t ww = warp; ww < BS_WORDS && wb + ww < n_rowwords; ww += BS_WARPS) {
            const uint64_t row0 = (wb + ww) * 32;  // relative row of the warp's first row
            const uint32_t rows_valid = (uint32_t)min((uint64_t)32, row_end - row_begin - row0);
            uint32_t *dst = sliced + row0 * ls + lg0;
            if (cols == BS_COLS) bitslice_warp<true>(in, out, ww, cols, lane, tl, dst, ls, rows_valid);
            else bitslice_warp<false>(in, out, ww, cols, lane, tl, dst, ls, rows_valid);
        }
    }
}

// Sliced probe.  CTA (x = block of QS_WARPS sub-chunks of positions, y = group of 32 column words = 1024 leaves,
// z = query).  A warp walks one sub-chunk of the query's (ascending) positions that fall into [row_begin, row_end):
// per position the 32 lanes read one 128-byte row segment (coalesced), and each lane adds its 32 one-bit leaf
// columns into 12 vertical (bit-plane) counters -- seven rows are first compressed 7 -> 3 planes with four
// LOP3 full adders, then rippled into the planes (32 LOP3 per 7 rows per 32 leaves).  Planes are unpacked once
// per sub-chunk into shared-memory leaf counters, and once per CTA into hits[] with global atomics.
// HBM traffic = one row segment per (position, leaf group): the algorithmic nP * ceil(L/8) bytes of SURVEY 8d.
constexpr int QS_WARPS = 8;
constexpr uint32_t QS_SUB = 4064;  // positions per warp sub-chunk: a multiple of 32, at most 4,095 (12 counter planes)

__device__ __forceinline__ void full_add(uint32_t a, uint32_t b, uint32_t c, uint32_t &sum, uint32_t &carry) {
    uint32_t s_, c_;  // the outputs may alias the inputs
    asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(s_) : "r"(a), "r"(b), "r"(c));  // a ^ b ^ c
    asm("lop3.b32 %0, %1, %2, %3, 0xE8;" : "=r"(c_) : "r"(a), "r"(b), "r"(c));  // majority
    sum = s_;
    carry = c_;
}

// W = u32 words per lane and load (1, 2 or 4): a warp reads 128 * W contiguous bytes of a row per position, and a CTA
// covers 1024 * W leaves.  Wider loads mean fewer, longer DRAM bursts per row (rows are 32-byte aligned: the stride is
// a multiple of 8 words) at the price of W times the counter registers.
template <int W>
struct QsVec;
template <>
struct QsVec<1> {
    uint32_t v[1];
    __device__ __forceinline__ void load(const uint32_t *p) { v[0] = __ldg(p); }
};
template <>
struct QsVec<2> {
    uint32_t v[2];
    __device__ __forceinline__ void load(const uint32_t *p) { const uint2 t = __ldg((const uint2 *)p); v[0] = t.x; v[1] = t.y; }
};
template <>
struct QsVec<4> {
    uint32_t v[4];
    __device__ __forceinline__ void load(const uint32_t *p) { const uint4 t = __ldg((const uint4 *)p); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
};

template <int W, int G, bool QFAST>
__global__ void __launch_bounds__(QS_WARPS * 32, (W == 2 && G == 2 ? 3 : 0))
query_sliced_kernel(const uint32_t *__restrict__ sliced, uint32_t n_leaves, uint32_t lw, uint32_t ls, uint64_t row_begin, uint64_t row_end,
                    const uint32_t *__restrict__ positions, const ulonglong2 *__restrict__ q_rng,
                    uint32_t *__restrict__ hits) {
    __shared__ uint32_t sm_hits[1024 * W];
    __shared__ unsigned long long sm_range[2];
    // grid: x = query (or chunk), z = chunk (or query), see msbt_query_batch_sliced
    const uint32_t q = QFAST ? blockIdx.x : blockIdx.z, chunk = QFAST ? blockIdx.z : blockIdx.x;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t col = (blockIdx.y * 32 + lane) * W;  // first column word of this lane (ls is a multiple of 8: whole vectors)
    for (uint32_t i = threadIdx.x; i < 1024 * W; i += blockDim.x) sm_hits[i] = 0;
    if (threadIdx.x == 0) {
        // positions are ascending: the ones inside [row_begin, row_end) form one contiguous range
        const unsigned long long b = q_rng[q].x, e = q_rng[q].y;
        unsigned long long lo = b, hi = e;
        while (lo < hi) { const unsigned long long mid = (lo + hi) >> 1; if (positions[mid] < row_begin) lo = mid + 1; else hi = mid; }
        sm_range[0] = lo;
        hi = e;
        while (lo < hi) { const unsigned long long mid = (lo + hi) >> 1; if (positions[mid] < row_end) lo = mid + 1; else hi = mid; }
        sm_range[1] = lo;
    }
    __syncthreads();
    const unsigned long long rb = sm_range[0], re = sm_range[1];
    const unsigned long long sub_b = rb + ((unsigned long long)chunk * QS_WARPS + warp) * QS_SUB;
    if (sub_b < re) {
        const unsigned long long sub_e = min(sub_b + (unsigned long long)QS_SUB, re);
        const bool live = col < lw;  // padding words (lw .. ls) may be read by a live vector; their leaves are dropped below
        const uint32_t *__restrict__ colp = sliced + (live ? col : 0);
        uint32_t P[W][12];
#pragma unroll
        for (int w = 0; w < W; w++)
#pragma unroll
            for (int i = 0; i < 12; i++) P[w][i] = 0;
        // 32 positions per round (one per lane; loaded one round ahead, so that their latency is not in front of every
        // round's row reads), four groups of 8 independent row reads, G groups in flight.  Counting is a carry-save chain
        // whose accumulators ARE the low counter planes: 8 rows fold into planes 0-2 with 7 full adders and leave one
        // carry of weight 8; two such carries meet in one more full adder on plane 3, and only its carry (weight 16)
        // ripples through planes 4-11: 2.9 ALU operations per row and lane word (the first version compressed 7 rows to 3
        // planes and rippled all 12 planes: 4.6).
        uint32_t nextpos = (sub_b + lane < sub_e) ? positions[sub_b + lane] : 0xFFFFFFFFu;
        for (unsigned long long base = sub_b; base < sub_e; base += 32) {
            const uint32_t mypos = nextpos;
            const unsigned long long idx = base + 32 + lane;
            nextpos = (idx < sub_e) ? positions[idx] : 0xFFFFFFFFu;
            uint32_t pend[W];
#pragma unroll
            for (int g8 = 0; g8 < 4; g8 += G) {
                QsVec<W> x[G][8];
#pragma unroll
                for (int g = 0; g < G; g++)
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        const uint32_t p = __shfl_sync(0xffffffffu, mypos, (g8 + g) * 8 + j);
                        if (p != 0xFFFFFFFFu && live) x[g][j].load(colp + ((uint64_t)p - row_begin) * ls);
                        else {
#pragma unroll
                            for (int w = 0; w < W; w++) x[g][j].v[w] = 0u;
                        }
                    }
#pragma unroll
                for (int g = 0; g < G; g++)
#pragma unroll
                    for (int w = 0; w < W; w++) {
                        uint32_t a, b, c, d, e, f, c8;
                        full_add(P[w][0], x[g][0].v[w], x[g][1].v[w], P[w][0], a);
                        full_add(P[w][0], x[g][2].v[w], x[g][3].v[w], P[w][0], b);
                        full_add(P[w][1], a, b, P[w][1], c);
                        full_add(P[w][0], x[g][4].v[w], x[g][5].v[w], P[w][0], d);
                        full_add(P[w][0], x[g][6].v[w], x[g][7].v[w], P[w][0], e);
                        full_add(P[w][1], d, e, P[w][1], f);
                        full_add(P[w][2], c, f, P[w][2], c8);  // carry of weight 8
                        if (((g8 + g) & 1) == 0) {
                            pend[w] = c8;
                        } else {
                            uint32_t k;
                            full_add(P[w][3], pend[w], c8, P[w][3], k);  // carry of weight 16
#pragma unroll
                            for (int i = 4; i < 12; i++) { const uint32_t t = P[w][i] & k; P[w][i] ^= k; k = t; }
                        }
                    }
            }
        }
        if (live) {
#pragma unroll
            for (int w = 0; w < W; w++) {
#pragma unroll 4
                for (int l = 0; l < 32; l++) {
                    uint32_t c = 0;
#pragma unroll
                    for (int i = 0; i < 12; i++) c |= ((P[w][i] >> l) & 1u) << i;
                    if (c) atomicAdd(sm_hits + (lane * W + w) * 32 + l, c);
                }
            }
        }
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < 1024 * W; i += blockDim.x) {
        const uint32_t leaf = blockIdx.y * 1024 * W + i;
        const uint32_t c = sm_hits[i];
        if (c && leaf < n_leaves) atomicAdd(hits + (uint64_t)q * n_leaves + leaf, c);
    }
}

// Row-coherent variant of the sliced probe for rows of several 256-byte pieces (more than 2,048 leaves).  In the kernel above
// the pieces (column groups) of a row are fetched by DIFFERENT CTAs (grid.y) at unrelated times; here the warps of one CTA walk the
// SAME positions, warp w reading piece w of every row, so the 1,280 bytes of a 10,000-leaf row are requested together.  A
// micro-benchmark of the bare access pattern (tools/microbench4.cu: random 1,280-byte rows of a 43 GB matrix read as five 256-byte
// warp loads) gives 2.07 TB/s when the five pieces of a row are requested at the same time against 1.37 TB/s when every warp
// reads a piece of its own random row: DRAM page locality.  Counters: 12 vertical planes per lane word fed by a carry-save chain
// (see the loop); a warp owns its column group, so the planes are unpacked straight into global atomics per sub-chunk.
constexpr int QSR_MAX_WARPS = 8;  // pieces of a row one CTA covers (10,000 leaves: 5); more leaves take grid.z
// (register-capped instantiations for 7 / 8 resident CTAs per SM were measured: 56 registers 82.4 ms, 48 registers with
// spills 92.0 ms, against 79.4 ms for the plain 63-register build with 6 CTAs per SM)
// (the chain with pending weight-16 / weight-32 carries needs 72 registers uncapped: probe 145.7 -> 143.6 ms per 512 MAGs;
// capped at 64 with 8 bytes of spill it keeps 6 CTAs per SM: 139.6 ms)
template <int G>   // groups of 8 row loads in flight per lane (1 or 2)
__global__ void __launch_bounds__(QSR_MAX_WARPS * 32, 4)   // 64 registers: 6 CTAs of 5 warps per SM at 10,000 leaves
query_sliced_rows_kernel(const uint32_t *__restrict__ sliced, uint32_t n_leaves, uint32_t lw, uint32_t ls, uint64_t row_begin, uint64_t row_end,
                         const uint32_t *__restrict__ positions, const ulonglong2 *__restrict__ q_rng, uint32_t *__restrict__ hits,
                         uint32_t n_groups, uint32_t sub) {
    constexpr int W = 2;
    __shared__ unsigned long long sm_range[2];
    const uint32_t q = blockIdx.x, chunk = blockIdx.y;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t group = blockIdx.z * QSR_MAX_WARPS + warp;     // column group = 64 words = 2,048 leaves
    if (threadIdx.x == 0) {
        const unsigned long long b = q_rng[q].x, e = q_rng[q].y;
        unsigned long long lo = b, hi = e;
        while (lo < hi) { const unsigned long long mid = (lo + hi) >> 1; if (positions[mid] < row_begin) lo = mid + 1; else hi = mid; }
        sm_range[0] = lo;
        hi = e;
        while (lo < hi) { const unsigned long long mid = (lo + hi) >> 1; if (positions[mid] < row_end) lo = mid + 1; else hi = mid; }
        sm_range[1] = lo;
    }
    __syncthreads();
    const unsigned long long rb = sm_range[0], re = sm_range[1];
    const unsigned long long sub_b = rb + (unsigned long long)chunk * sub;   // sub <= QS_SUB positions: 12 counter planes
    if (sub_b >= re || group >= n_groups) return;
    const unsigned long long sub_e = min(sub_b + (unsigned long long)sub, re);
    const uint32_t col = (group * 32 + lane) * W;
    const bool live = col < lw;
    const uint32_t *__restrict__ colp = sliced + (live ? col : 0);
    uint32_t P[W][12];
#pragma unroll
    for (int w = 0; w < W; w++)
#pragma unroll
        for (int i = 0; i < 12; i++) P[w][i] = 0;
    // 32 positions per round (one per lane), four groups of 8 rows.  Counting is a carry-save chain whose accumulators ARE
    // the low counter planes: 8 rows fold into planes 0-2 with 7 full adders and leave one carry of weight 8; two such
    // carries meet in one more full adder on plane 3, and only its carry (weight 16) ripples through planes 4-11 -- 2.9 ALU
    // operations per row and lane word instead of 4.6 for the 7 -> 3 compression + 12-plane ripple of the kernel above.
    // The weight-16 carries of a round's two halves meet on plane 4, and the weight-32 carries of two consecutive rounds on
    // plane 5; only that carry (weight 64, once per 64 rows) ripples through planes 6-11: 2.2 operations per row and word.
    uint32_t nextpos = (sub_b + lane < sub_e) ? positions[sub_b + lane] : 0xFFFFFFFFu;
    uint32_t pend32[W];
    bool odd_round = false;
    for (unsigned long long base = sub_b; base < sub_e; base += 32) {
        const uint32_t mypos = nextpos;
        const unsigned long long idx = base + 32 + lane;
        nextpos = (idx < sub_e) ? positions[idx] : 0xFFFFFFFFu;
        uint32_t pend[W], pend16[W];
#pragma unroll
        for (int g8 = 0; g8 < 4; g8 += G) {
            QsVec<W> x[G][8];
#pragma unroll
            for (int g = 0; g < G; g++)
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const uint32_t p = __shfl_sync(0xffffffffu, mypos, (g8 + g) * 8 + j);
                    if (p != 0xFFFFFFFFu && live) x[g][j].load(colp + ((uint64_t)p - row_begin) * ls);
                    else {
#pragma unroll
                        for (int w = 0; w < W; w++) x[g][j].v[w] = 0u;
                    }
                }
#pragma unroll
            for (int g = 0; g < G; g++)
#pragma unroll
                for (int w = 0; w < W; w++) {
                    uint32_t a, b, c, d, e, f, c8;
                    full_add(P[w][0], x[g][0].v[w], x[g][1].v[w], P[w][0], a);
                    full_add(P[w][0], x[g][2].v[w], x[g][3].v[w], P[w][0], b);
                    full_add(P[w][1], a, b, P[w][1], c);
                    full_add(P[w][0], x[g][4].v[w], x[g][5].v[w], P[w][0], d);
                    full_add(P[w][0], x[g][6].v[w], x[g][7].v[w], P[w][0], e);
                    full_add(P[w][1], d, e, P[w][1], f);
                    full_add(P[w][2], c, f, P[w][2], c8);      // carry of weight 8
                    if (((g8 + g) & 1) == 0) {
                        pend[w] = c8;
                    } else {
                        uint32_t k;
                        full_add(P[w][3], pend[w], c8, P[w][3], k);   // k: carry of weight 16
                        if (((g8 + g) & 2) == 0) {
                            pend16[w] = k;
                        } else {
                            uint32_t k32;
                            full_add(P[w][4], pend16[w], k, P[w][4], k32);   // carry of weight 32, one per round
                            if (!odd_round) {
                                pend32[w] = k32;
                            } else {
                                uint32_t k64;
                                full_add(P[w][5], pend32[w], k32, P[w][5], k64);
#pragma unroll
                                for (int i = 6; i < 12; i++) { const uint32_t t = P[w][i] & k64; P[w][i] ^= k64; k64 = t; }
                            }
                        }
                    }
                }
        }
        odd_round = !odd_round;
    }
    if (odd_round) {  // an unpaired weight-32 carry is still pending
#pragma unroll
        for (int w = 0; w < W; w++) {
            uint32_t k = pend32[w];
#pragma unroll
            for (int i = 5; i < 12; i++) { const uint32_t t = P[w][i] & k; P[w][i] ^= k; k = t; }
        }
    }
    if (live) {
        uint32_t *__restrict__ hq = hits + (uint64_t)q * n_leaves;
#pragma unroll
        for (int w = 0; w < W; w++) {
            const uint32_t leaf0 = (col + w) * 32;
#pragma unroll 4
            for (int l = 0; l < 32; l++) {
                uint32_t c = 0;
#pragma unroll
                for (int i = 0; i < 12; i++) c |= ((P[w][i] >> l) & 1u) << i;
                if (c && leaf0 + l < n_leaves) atomicAdd(hq + leaf0 + l, c);
            }
        }
    }
}

// Row stride of the bit-sliced matrix: ceil(n_leaves / 32) words rounded up to 8 words, so that every row starts on a
// 32-byte sector boundary (10,000 leaves: 313 -> 320 words; an unaligned 128-byte row segment costs five sectors, not four).
extern "C" uint32_t msbt_sliced_row_words(uint32_t n_leaves) { return (((n_leaves + 31) / 32) + 7u) & ~7u; }

extern "C" int msbt_bitslice_build(msbt_ctx *ctx, const uint64_t *const *h_leaves, uint32_t n_leaves,
                                   uint64_t row_begin, uint64_t row_end, uint32_t *d_sliced, void *stream_) {
    if (!ctx) return MSBT_EINVAL;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n_leaves == 0 || row_end <= row_begin) return MSBT_OK;
    if (!d_sliced) return fail(ctx, MSBT_EINVAL, "bitslice_build: null output");
    if (row_begin & 31) return fail(ctx, MSBT_EINVAL, "bitslice_build: row_begin must be a multiple of 32");
    int rc = check_ptrs(ctx, h_leaves, n_leaves);
    if (rc) return rc;
    CTX_DEVICE(ctx);
    const uint64_t *const *d_l;
    rc = upload_ptrs(ctx, h_leaves, n_leaves, stream, &d_l);
    if (rc) return rc;
    const uint32_t lw = (n_leaves + 31) / 32;
    if ((lw + BS_COLS - 1) / BS_COLS > 65535) return fail(ctx, MSBT_EINVAL, "bitslice_build: too many leaves");
    const uint64_t n_rowwords = (row_end - row_begin + 31) >> 5;
    CU_TRY(ctx, cudaFuncSetAttribute(bitslice_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BS_SMEM_BYTES));
    const uint32_t gy = (lw + BS_COLS - 1) / BS_COLS;
    const uint64_t gx = std::max<uint64_t>(1, std::min<uint64_t>((n_rowwords + BS_WORDS - 1) / BS_WORDS,
                                                                  std::max<uint64_t>(1, (uint64_t)ctx->num_sms * 8 / gy)));
    bitslice_kernel<<<dim3((unsigned)gx, gy), BS_THREADS, BS_SMEM_BYTES, stream>>>(d_l, n_leaves, row_begin, row_end, lw,
                                                                             msbt_sliced_row_words(n_leaves), d_sliced);
    LAUNCH_CHECK(ctx);
    return MSBT_OK;
}

extern "C" int msbt_query_batch_sliced_ranges(msbt_ctx *ctx, const uint32_t *d_sliced, uint32_t n_leaves, uint64_t row_begin,
                                              uint64_t row_end, const uint32_t *d_positions, const uint64_t *h_query_ranges,
                                              uint32_t n_queries, uint32_t *d_hits, void *stream_) {
    if (!ctx) return MSBT_EINVAL;
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n_queries == 0 || n_leaves == 0) return MSBT_OK;
    if (!d_sliced || !h_query_ranges || !d_hits) return fail(ctx, MSBT_EINVAL, "query_batch_sliced: null argument");
    if (n_queries > 65535) return fail(ctx, MSBT_EINVAL, "query_batch_sliced: at most 65535 queries per call");
    uint64_t max_pos = 0;
    for (uint32_t q = 0; q < n_queries; q++) {
        if (h_query_ranges[2 * q + 1] < h_query_ranges[2 * q]) return fail(ctx, MSBT_EINVAL, "query %u: range end before begin", q);
        max_pos = std::max<uint64_t>(max_pos, h_query_ranges[2 * q + 1] - h_query_ranges[2 * q]);
    }
    CTX_DEVICE(ctx);
    // hits accumulate with atomics: clear first (partial results of other bit ranges are summed by the caller)
    CU_TRY(ctx, cudaMemsetAsync(d_hits, 0, (size_t)n_queries * n_leaves * 4, stream));
    if (max_pos == 0) return MSBT_OK;
    if (!d_positions) return fail(ctx, MSBT_EINVAL, "query_batch_sliced: null positions");
    const size_t rng_b = (size_t)n_queries * 16;
    int rc = ensure_scratch(ctx, rng_b);
    if (rc) return rc;
    rc = stage_begin(ctx, rng_b);
    if (rc) return rc;
    const ulonglong2 *d_rng;
    rc = upload_ranges(ctx, h_query_ranges, n_queries, 0, stream, &d_rng);
    if (rc) return rc;
    rc = stage_end(ctx, stream);
    if (rc) return rc;
    const uint32_t lw = (n_leaves + 31) / 32;
    const uint64_t per_cta = (uint64_t)QS_WARPS * QS_SUB;
    const uint32_t gx = (uint32_t)((max_pos + per_cta - 1) / per_cta);
    const uint32_t ls = msbt_sliced_row_words(n_leaves);
    // words per lane: 8-byte loads (256 contiguous bytes of a row per warp) once a row has at least 64 words; 16-byte loads
    // need 134 registers and lose to the occupancy drop (tools/bench_query.py, 10,000 leaves: W = 1 / 2 / 4 ->
    // 48 % / 66 % / 40 % of the HBM peak)
    int W = ls >= 64 ? 2 : 1;
    if (const char *e = getenv("MSBT_QS_W")) { const int w = atoi(e); if (w == 1 || w == 2 || w == 4) W = w; }  // tuning aid
    const uint32_t gy = (ls + 32 * W - 1) / (32 * W);
    if (gy > 65535) return fail(ctx, MSBT_EINVAL, "query_batch_sliced: too many leaves");
    // Block order.  The queries of a batch walk the same ascending row space, and chunk c of every query covers about the
    // same row range (positions are uniform hashes).  With the QUERY as the fastest grid dimension the CTAs that run
    // together probe the same rows, so a row that several queries hit (256 MAGs against 2^25-row filters: 15 per row)
    // comes from L2 instead of DRAM.  MSBT_QS_ORDER=0 restores chunk-fastest order (one query after the other).
    // rows of several 256-byte pieces: the row-coherent kernel (MSBT_QS_ROWS=0 keeps the column-group kernel)
    bool rows_kernel = W == 2 && gy >= 2;
    if (const char *e = getenv("MSBT_QS_ROWS")) rows_kernel = rows_kernel && atoi(e) != 0;
    if (rows_kernel) {
        uint32_t sub = QS_SUB;
        if (const char *e = getenv("MSBT_QSR_SUB")) { const int v = atoi(e); if (v >= 32 && v <= 4064) sub = (uint32_t)v / 32 * 32; }  // tuning aid
        const uint64_t chunks = (max_pos + sub - 1) / sub;
        if (chunks <= 65535) {
            const uint32_t warps = std::min<uint32_t>(gy, QSR_MAX_WARPS);
            const dim3 grid_r(n_queries, (unsigned)chunks, (gy + QSR_MAX_WARPS - 1) / QSR_MAX_WARPS);
            int g_rows = 1;
            if (const char *e = getenv("MSBT_QSR_G")) g_rows = atoi(e);  // tuning aid
            if (g_rows == 2)
                query_sliced_rows_kernel<2><<<grid_r, warps * 32, 0, stream>>>(d_sliced, n_leaves, lw, ls, row_begin, row_end, d_positions, d_rng, d_hits, gy, sub);
            else
                query_sliced_rows_kernel<1><<<grid_r, warps * 32, 0, stream>>>(d_sliced, n_leaves, lw, ls, row_begin, row_end, d_positions, d_rng, d_hits, gy, sub);
            LAUNCH_CHECK(ctx);
            return MSBT_OK;
        }
    }
    bool qfast = n_queries > 1 && gx <= 65535;
    if (const char *e = getenv("MSBT_QS_ORDER")) qfast = qfast && atoi(e) != 0;
    const dim3 grid = qfast ? dim3(n_queries, gy, gx) : dim3(gx, gy, n_queries);
    int G = W == 1 ? 2 : 1;  // groups of 7 row reads in flight per lane (W = 1, 1,024 leaves: 46.5 % -> 47.6 % of the HBM peak)
    if (const char *e = getenv("MSBT_QS_G")) { const int g = atoi(e); if (g == 1 || g == 2) G = g; }  // tuning aid
#define MSBT_QS_LAUNCH2(WW, GG, QF) query_sliced_kernel<WW, GG, QF><<<grid, QS_WARPS * 32, 0, stream>>>(d_sliced, n_leaves, lw, ls, row_begin, row_end, d_positions, d_rng, d_hits)
#define MSBT_QS_LAUNCH(WW, GG) do { if (qfast) MSBT_QS_LAUNCH2(WW, GG, true); else MSBT_QS_LAUNCH2(WW, GG, false); } while (0)
    if (W == 4) MSBT_QS_LAUNCH(4, 1);
    else if (W == 2 && G == 2) MSBT_QS_LAUNCH(2, 2);
    else if (W == 2) MSBT_QS_LAUNCH(2, 1);
    else if (G == 2) MSBT_QS_LAUNCH(1, 2);
    else MSBT_QS_LAUNCH(1, 1);
#undef MSBT_QS_LAUNCH2
#undef MSBT_QS_LAUNCH
    LAUNCH_CHECK(ctx);
    return MSBT_OK;
}

extern "C" int msbt_query_batch_sliced(msbt_ctx *ctx, const uint32_t *d_sliced, uint32_t n_leaves, uint64_t row_begin,
                                       uint64_t row_end, const uint32_t *d_positions, const uint64_t *h_query_offsets,
                                       uint32_t n_queries, uint32_t *d_hits, void *stream_) {
    if (!ctx) return MSBT_EINVAL;
    if (n_queries == 0 || n_leaves == 0) return MSBT_OK;
    std::vector<uint64_t> rng;
    int rc = ranges_from_offsets(ctx, h_query_offsets, n_queries, rng);
    if (rc) return rc;
    return msbt_query_batch_sliced_ranges(ctx, d_sliced, n_leaves, row_begin, row_end, d_positions, rng.data(), n_queries, d_hits, stream_);
}
