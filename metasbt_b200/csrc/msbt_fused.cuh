// msbt_fused.cuh -- K1 variant 3 ("fused"): ONE persistent, warp-specialised kernel per batch that
// replaces `howdesbt makebf` (/root/reference/metasbt/core.py:3920-3931) for all genomes of the batch.
//
// One CTA per SM, 1024 threads, two roles that run concurrently on the same SM:
//
//   P  (20 warps)  partition: claim a chunk of FU_CHUNK_BASES k-mer starts, hash every canonical k-mer,
//                  stage its position in shared memory by bucket (bucket = pos >> bucket_shift; at most NBM
//                  buckets per filter, NBM = 512 below 2^30 bits and 2048 from there, see FuGeom), then flush
//                  each staged row with one global reservation and aligned 16-byte stores into an L2-resident
//                  scratch ring (one slot per genome in flight).
//                  This role is instruction-issue bound (64-bit MurmurHash3 on a 32-bit ALU).
//   B  (12 warps)  build: claim a (genome, bucket) item once all chunks of that genome are flushed, pull
//                  the bucket's positions into registers, and for each <= 64 KiB tile of the bucket: zero
//                  the tile in shared memory, OR the positions into it (red.shared.or), and write the finished
//                  tile to HBM with one TMA bulk store (cp.async.bulk.global.shared::cta).  This role is
//                  HBM-write bound and mostly waits; P fills the issue slots it leaves idle.
//
// The filter is written exactly once (zeros included): no memset, no read-modify-write in HBM.  HBM traffic
// per genome = packed input + filter; everything else stays in shared memory or the L2-resident ring.
//
// Flow control is through global counters (all zeroed by one memset before the launch):
//   sync[0] next chunk to claim, sync[1] next item to claim,
//   done[g]      P warps that flushed a chunk of genome g (B waits for done[g] == n_chunks(g) * P warps)
//   consumed[g]  buckets of genome g built       (P waits for consumed[g - ring] == NB before reusing a slot)
//   dirty[g]     set when a bucket of the ring overflowed (skewed inputs): genome g is redone afterwards by
//                the direct kernel, which ORs every bit into the already written filter.  The reservation that
//                did not fit pads the tail it leaves unwritten with a position of the same bucket, because B
//                reads min(count, capacity) entries and the redo pass can only add bits.
// Claims are handed out in genome order, so a waiting role only ever waits on strictly earlier work that
// is already owned by a running CTA: there is no circular wait, and a CTA that starts late simply joins.
#pragma once

#ifndef MSBT_FU_P_THREADS
#define MSBT_FU_P_THREADS 640
#endif
constexpr int FU_P_THREADS = MSBT_FU_P_THREADS;      // partition warps (role P)
constexpr int FU_THREADS = 1024;
constexpr int FU_B_THREADS = FU_THREADS - FU_P_THREADS;  // tile-build warps (role B)
constexpr uint32_t FU_CHUNK_BASES = FU_P_THREADS * 32;  // k-mer starts per chunk
#ifndef MSBT_FU_NB
#define MSBT_FU_NB 2048
#endif
// The bucket count is a template parameter of the kernel (NBM, the maximum number of buckets per filter): 2048 for filters
// of 2^30 bits and more (a bucket is one 64 KiB tile: single-pass build), 512 below (fewer, larger buckets: a quarter of the
// per-tile chains on the B side; 5 Mbp genomes, 2^22 .. 2^28 bits: 34.5 -> 27.4 us/genome, 2^29: 37.7 -> 30.8;
// profiles/r02_k1_bucket_geometry.json).  Everything that depends on it is FuGeom<NBM>.
constexpr uint32_t FU_NBM_LARGE = MSBT_FU_NB;
constexpr uint32_t FU_NBM_SMALL = 512;
constexpr uint32_t FU_TILE_LOG2 = 19;                   // <= 2^19 bits = 64 KiB per tile
constexpr uint32_t FU_TILE_BYTES = 1u << (FU_TILE_LOG2 - 3);
template <uint32_t NBM>
struct FuGeom {
    static constexpr uint32_t MAX_BUCKETS = NBM;
    static constexpr uint32_t STAGE_WORDS = (32768 / NBM) * (NBM + 1);  // staging rows: (NB + 1) x cap
    // 16-byte groups of bucket positions per B thread in registers
    static constexpr int REG_QUADS = (768 * (2048 / (int)NBM) + FU_B_THREADS - 1) / FU_B_THREADS;
    static constexpr int REG_ENTRIES = 4 * REG_QUADS;
    // dynamic shared memory: [tile 64 KiB][cnt NB][stage][claims]
    static constexpr size_t SMEM_BYTES = FU_TILE_BYTES + NBM * 4 + (size_t)STAGE_WORDS * 4 + 64;
};
constexpr uint32_t FU_SKIP = 0xFFFFFFFFu;

#ifdef MSBT_FU_TRACE
// Kernel-tuning builds only: per-phase clock64() totals of CTA 0 (P phases 0..7, B phases 8..15), read back with
// msbt_debug_trace().  Not compiled into the shipped library.
__device__ unsigned long long fu_trace[24];
#define FU_T(var) const long long var = clock64();
#define FU_ACC(idx, a, b) if (blockIdx.x == 0) atomicAdd(&fu_trace[idx], (unsigned long long)((b) - (a)));
#else
#define FU_T(var)
#define FU_ACC(idx, a, b)
#endif

struct FusedParams {
    const uint64_t *packed;
    const uint32_t *run_ends;
    const DevGenome *genomes;
    uint32_t n_genomes;
    uint32_t n_chunks_total;
    uint32_t bucket_shift;   // bucket = pos >> bucket_shift
    uint32_t NB;             // buckets per filter (<= NBM)
    uint32_t cap;            // staging entries per bucket per chunk (multiple of 4)
    uint32_t quads;          // cap / 4
    uint32_t lanes_per_row;  // flush: power of two >= min(quads, 32)
    uint32_t quad_magic;     // ceil(2^32 / quads): idx / quads == umulhi(idx, quad_magic) for idx < NB * quads
    uint32_t tile_log2;      // min(bucket_shift, FU_TILE_LOG2)
    uint32_t tiles_per_bucket_log2;
    uint32_t gcap;           // ring entries per bucket (multiple of 4)
    uint32_t ring;           // genomes in flight
    uint32_t fast_b;         // 1: bucket == one full 64 KiB tile, every destination 16-byte aligned (the lean B loop applies)
    uint32_t nb_magic;       // ceil(2^32 / NB): item / NB == umulhi(item, nb_magic) for item < n_items
    uint64_t words;          // u64 words per filter
    uint32_t *sync;          // [0] next chunk, [1] next item
    uint32_t *done;          // [n_genomes]
    uint32_t *consumed;      // [n_genomes]
    uint32_t *dirty;         // [n_genomes]
    uint32_t *gcnt;          // [n_genomes][NB] entries appended per (genome, bucket)
    uint32_t *gbuf;          // [ring][NB][gcap]
    uint64_t *filters;
};

__device__ __forceinline__ void fu_bar(uint32_t id, uint32_t n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ uint32_t fu_ld_acquire(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// Poll with relaxed loads (an acquire load invalidates the SM's L1 every time), acquire once at the end.
__device__ __forceinline__ void fu_wait_ge(const uint32_t *p, uint32_t want) {
    if (fu_ld_acquire(p) >= want) return;
    uint32_t v;
    do {
        __nanosleep(200);
        asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    } while (v < want);
    (void)fu_ld_acquire(p);
}

// Predicated shared-memory OR: a rejected position costs a predicate, not a divergent branch.
__device__ __forceinline__ void reds_or_if(uint32_t saddr, uint32_t v, bool ok) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p red.shared.or.b32 [%0], %1;\n\t}" ::"r"(saddr), "r"(v), "r"((uint32_t)ok) : "memory");
}

// AV: all 32 k-mer starts of the thread's word are valid (the common case: no run boundary in the word), and with
// FULL (modulus <= num_bits) every position is in range: the k-mer needs no predicate at all.
#define FU_HASH(J)                                                                                                   \
    {                                                                                                                \
        uint64_t lo_, hi_;                                                                                           \
        kb.template get<(H * 4 + J)>(lo_, hi_);                                                                      \
        const uint64_t p_ = msbt_position_v<POW2>(msbt_kmer_hash_v<HV>(lo_, hi_, len, spec.pos.seed), spec.pos);     \
        pos4[J] = (uint32_t)p_;                                                                                      \
        ok4[J] = (AV || ((vg >> (H * 4 + J)) & 1u)) && (FULL || p_ < spec.pos.num_bits);                             \
        b4[J] = ok4[J] ? (pos4[J] >> P.bucket_shift) : P.NB;                                                         \
    }
#define FU_RANK(J) r4[J] = (AV && FULL) ? atoms_inc(cnt_s + b4[J] * 4u) : atoms_inc_if(cnt_s + b4[J] * 4u, ok4[J]);
#define FU_STORE(J)                                                                   \
    sts_if(stage_s + (b4[J] * P.cap + r4[J]) * 4u, pos4[J], r4[J] < P.cap);           \
    over |= (r4[J] >= P.cap);
// staging row full (about 0.1 % of k-mers at lambda = 24, cap = 36): append straight to the ring bucket
#define FU_COLD(J)                                                                    \
    if (r4[J] >= P.cap) fu_append_cold(gc, gb, P.gcap, b4[J], pos4[J], dirty_g);

__device__ __forceinline__ uint32_t atoms_inc(uint32_t saddr) {
    uint32_t r;
    asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(r) : "r"(saddr) : "memory");
    return r;
}

__device__ __noinline__ void fu_append_cold(uint32_t *gc, uint32_t *gb, uint32_t gcap, uint32_t b, uint32_t pos,
                                            uint32_t *dirty_g) {
    const uint32_t o = atomicAdd(gc + b, 4u);
    if (o + 4u <= gcap) {
        *(uint4 *)(gb + (uint64_t)b * gcap + o) = make_uint4(pos, pos, pos, pos);
    } else {
        *dirty_g = 1u;  // the direct kernel redoes this genome after the fused kernel
    }
}

// One half step = 4 k-mers: four independent hash chains, then the four rank atomics, then the four staged stores.
// Measured alternatives that did not help: software-pipelining the stores one half step behind the atomics (12 more
// live registers spill at the 64-register budget of a 1024-thread CTA), and eight chains per step with registers
// moved from the B to the P warps by setmaxnreg (-DMSBT_FU_SETMAXNREG): the loop is already issue/IMAD-pipe bound.
template <int KW, int HV, bool POW2, bool FULL, bool AV, int H>
__device__ __forceinline__ void fu_half_step(const KmerBlock<KW> &kb, uint32_t vg, const KmerSpec &spec, const FusedParams &P,
                                             uint32_t len, uint32_t cnt_s, uint32_t stage_s, uint32_t *gc, uint32_t *gb,
                                             uint32_t *dirty_g) {
    uint32_t pos4[4], b4[4], r4[4];
    bool ok4[4];
    FU_HASH(0) FU_HASH(1) FU_HASH(2) FU_HASH(3)
    FU_RANK(0) FU_RANK(1) FU_RANK(2) FU_RANK(3)
    bool over = false;
    FU_STORE(0) FU_STORE(1) FU_STORE(2) FU_STORE(3)
    if (over) { FU_COLD(0) FU_COLD(1) FU_COLD(2) FU_COLD(3) }
}
// the 32 k-mers of one packed word
template <int KW, int HV, bool POW2, bool FULL, bool AV>
__device__ __forceinline__ void fu_word(KmerBlock<KW> &kb, uint32_t vm, const KmerSpec &spec, const FusedParams &P, uint32_t len,
                                        uint32_t cnt_s, uint32_t stage_s, uint32_t *gc, uint32_t *gb, uint32_t *dirty_g) {
#pragma unroll 1
    for (uint32_t grp = 0; grp < 4; grp++) {
        const uint32_t vg = AV ? 0xFFu : ((vm >> (8 * grp)) & 0xFFu);
        fu_half_step<KW, HV, POW2, FULL, AV, 0>(kb, vg, spec, P, len, cnt_s, stage_s, gc, gb, dirty_g);
        fu_half_step<KW, HV, POW2, FULL, AV, 1>(kb, vg, spec, P, len, cnt_s, stage_s, gc, gb, dirty_g);
        kb.advance8();
    }
}

template <int KW, int HV, bool POW2, bool FULL, uint32_t NBM>
__global__ void __launch_bounds__(FU_THREADS, 1)
kmer_bloom_fused_kernel(FusedParams P, KmerSpec spec) {
    constexpr uint32_t FU_MAX_BUCKETS = FuGeom<NBM>::MAX_BUCKETS;
    constexpr uint32_t FU_STAGE_WORDS = FuGeom<NBM>::STAGE_WORDS;
    constexpr int FU_REG_QUADS = FuGeom<NBM>::REG_QUADS;
    constexpr int FU_REG_ENTRIES = FuGeom<NBM>::REG_ENTRIES;
    extern __shared__ __align__(128) uint32_t fu_smem[];
    uint32_t *tile = fu_smem;                                        // B: one <= 64 KiB tile
    uint32_t *cnt = fu_smem + FU_TILE_BYTES / 4;                     // P: [NB + 1] staged per bucket (+ dummy)
    uint32_t *stage = cnt + FU_MAX_BUCKETS;                          // P: [NB + 1][cap], rows 16-byte aligned
    volatile uint32_t *claim = stage + FU_STAGE_WORDS;               // [0..1] P: chunk, genome; [4..5] B: item
    const uint32_t NB = P.NB;

    // MSBT_FU_B_FIRST puts the tile-build role on the low warp ids (a scheduler that breaks ties oldest-first would favour them)
#ifdef MSBT_FU_B_FIRST
    const bool role_p = threadIdx.x >= FU_B_THREADS;
    const uint32_t role_tid = role_p ? threadIdx.x - FU_B_THREADS : threadIdx.x;
#else
    const bool role_p = threadIdx.x < FU_P_THREADS;
    const uint32_t role_tid = role_p ? threadIdx.x : threadIdx.x - FU_P_THREADS;
#endif
    if (role_p) {
        // ------------------------------------------------------------------ P: partition
#ifdef MSBT_FU_SETMAXNREG
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(MSBT_FU_P_REGS));  // registers handed over by the B warps
#endif
        const uint32_t tid = role_tid;
        const uint32_t cnt_s = (uint32_t)__cvta_generic_to_shared(cnt);
        const uint32_t stage_s = (uint32_t)__cvta_generic_to_shared(stage);
        const uint32_t k = spec.k, len = (k + 3) >> 2;
        // tid 0 claims chunks two ahead (the atomic's round trip is never waited for) and resolves the next
        // chunk's genome / ring slot while the other warps flush, so a chunk starts without exposed latency.
        uint32_t pend = 0, nxt = 0, nxt_gi = 0, nxt_free = 0;
        if (tid == 0) {
            const uint32_t c = atomicAdd(P.sync + 0, 1u);
            pend = atomicAdd(P.sync + 0, 1u);
            uint32_t gi = 0;
            if (c < P.n_chunks_total) {
                uint32_t lo = 0, hi = P.n_genomes;  // last genome with fchunk_begin <= c
                while (hi - lo > 1) {
                    const uint32_t mid = (lo + hi) >> 1;
                    if (P.genomes[mid].fchunk_begin <= c) lo = mid; else hi = mid;
                }
                gi = lo;
                if (gi >= P.ring) fu_wait_ge(P.consumed + (gi - P.ring), NB);  // ring slot free again
            }
            claim[0] = c;
            claim[1] = gi;
        }
        for (uint32_t b = tid; b < NB; b += FU_P_THREADS) cnt[b] = 0;
        while (true) {
            FU_T(tp0)
            // barrier A: the claim is visible and every staging counter is zero again
            fu_bar(1, FU_P_THREADS);
            const uint32_t c = claim[0], gi = claim[1];
            if (c >= P.n_chunks_total) break;
            FU_T(tp1)
            const DevGenome g = P.genomes[gi];
            const uint32_t slot = gi % P.ring;
            uint32_t *gc = P.gcnt + (uint64_t)gi * NB;  // per genome: never reset, never reused
            uint32_t *gb = P.gbuf + (uint64_t)slot * NB * P.gcap;
            uint32_t *dirty_g = P.dirty + gi;

            const uint32_t w = (c - g.fchunk_begin) * FU_P_THREADS + tid;
            const uint32_t s0 = w * 32;
            uint32_t vm = 0;
            if (s0 < g.n_bases) vm = valid_mask(P.run_ends + g.run_off, g.n_runs, s0, k);
            if (vm == 0xFFFFFFFFu) {
                KmerBlock<KW> kb;
                kb.init(P.packed + g.packed_word_off + w, k);
                fu_word<KW, HV, POW2, FULL, true>(kb, vm, spec, P, len, cnt_s, stage_s, gc, gb, dirty_g);
            } else if (vm) {
                KmerBlock<KW> kb;
                kb.init(P.packed + g.packed_word_off + w, k);
                fu_word<KW, HV, POW2, FULL, false>(kb, vm, spec, P, len, cnt_s, stage_s, gc, gb, dirty_g);
            }
            fu_bar(1, FU_P_THREADS);  // barrier B: the chunk is staged
            FU_T(tp2)
            // Flush, warp by warp (each warp owns a contiguous range of staging rows, so no CTA barrier is needed):
            //  A  one lane per row: one ring reservation per (chunk, bucket), rounded up to whole quads; the result is
            //     packed back into the row's counter as (offset << 8) | n;
            //  B  `lpr` adjacent lanes per row copy its quads with aligned 16-byte loads/stores: a quarter warp reads
            //     consecutive quads of consecutive rows (bank-conflict free) and a row's quads leave in one instruction
            //     as one contiguous run (fewer L1 sector requests than one lane per row).  Pad slots repeat the row's
            //     first entry (OR is idempotent).  The counter is zeroed for the next chunk.
            {
                const uint32_t lane = tid & 31u, warp = tid >> 5;
                constexpr uint32_t P_WARPS = FU_P_THREADS / 32;
                const uint32_t rpw = (NB + P_WARPS - 1) / P_WARPS;
                const uint32_t r0 = warp * rpw, r1 = min(r0 + rpw, NB);
                // all reservations of a lane are issued before the first result is used (one L2 round trip, not four)
                for (uint32_t rb = r0; rb < r1; rb += 128) {
                    uint32_t nn[4], oo[4];
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const uint32_t b = rb + lane + 32 * j;
                        nn[j] = (b < r1) ? min(cnt[b], P.cap) : 0u;
                        oo[j] = nn[j] ? atomicAdd(gc + b, (nn[j] + 3u) & ~3u) : 0u;
                    }
                    if (rb == r0 && tid == 0) {
                        // tid 0 resolves the next chunk's claim while the reservations are in flight
                        nxt = pend;
                        pend = atomicAdd(P.sync + 0, 1u);
                        nxt_gi = gi;  // claims are monotonic: scan forward from the current genome
                        if (nxt < P.n_chunks_total) {
                            while (nxt_gi + 1 < P.n_genomes && P.genomes[nxt_gi + 1].fchunk_begin <= nxt) nxt_gi++;
                            // relaxed is enough: P only WRITES the slot after seeing the count, and B bumped it after its reads
                            if (nxt_gi >= P.ring) asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(nxt_free) : "l"(P.consumed + (nxt_gi - P.ring)) : "memory");
                            else nxt_free = NB;
                        }
                    }
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const uint32_t b = rb + lane + 32 * j;
                        if (b < r1) {
                            uint32_t n = nn[j];
                            if (n && oo[j] + ((n + 3u) & ~3u) > P.gcap) {
                                // ring bucket full (skewed input): the direct kernel redoes this genome afterwards.  B reads
                                // min(count, gcap) entries of the bucket, so the tail this reservation leaves unwritten is
                                // padded with a position that does belong to the bucket (OR is idempotent) -- stale ring
                                // content there would set bits the redo pass cannot take back.
                                *dirty_g = 1u;
                                const uint32_t first = stage[b * P.cap];
                                for (uint32_t o = oo[j]; o + 4u <= P.gcap; o += 4u)
                                    *(uint4 *)(gb + (uint64_t)b * P.gcap + o) = make_uint4(first, first, first, first);
                                n = 0;
                            }
                            cnt[b] = (oo[j] << 8) | n;
                        }
                    }
                }
                __syncwarp();
                FU_T(tf1)
                if (tid == 0) { FU_ACC(5, tp2, tf1) }
                const uint4 *stage4 = (const uint4 *)stage;
                uint4 *gb4 = (uint4 *)gb;
                const uint32_t lpr = P.lanes_per_row, rows_per_it = 32u / lpr;
                const uint32_t q0 = lane & (lpr - 1u), sub = lane / lpr;
                for (uint32_t rb = r0; rb < r1; rb += rows_per_it) {
                    const uint32_t b = rb + sub;
                    if (b < r1) {
                        const uint32_t info = cnt[b];
                        const uint32_t n = info & 0xFFu, o = info >> 8;
                        const uint32_t nq = (n + 3u) >> 2;
                        const uint4 *src4 = stage4 + b * P.quads;
                        uint4 *dst4 = gb4 + (((uint64_t)b * P.gcap + o) >> 2);
                        for (uint32_t q = q0; q < nq; q += lpr) {
                            uint4 v = src4[q];
                            if (q + 1 == nq) {
                                const uint32_t first = stage[b * P.cap];
                                const uint32_t have = n - q * 4u;
                                if (have < 2) v.y = first;
                                if (have < 3) v.z = first;
                                if (have < 4) v.w = first;
                            }
                            dst4[q] = v;
                        }
                    }
                    __syncwarp();
                    if (b < r1 && q0 == 0) cnt[b] = 0;
                }
            }
            FU_T(tf2)
            if (tid == 0) { FU_ACC(6, tp2, tf2) }
            // Publish per warp (done[g] counts warps, not chunks): no CTA barrier between the flush and the publish, and
            // the publish never waits behind tid 0's ring wait below (B needs it to free that very ring slot).
            // bar.warp.sync orders the lanes' ring writes before lane 0's fence, and the fence is cumulative: one
            // MEMBAR per warp instead of 32
            __syncwarp();
            if ((tid & 31u) == 0) {
                __threadfence();
                atomicAdd(P.done + gi, 1u);
            }
            FU_T(tp4)
            if (tid == 0) {
                if (nxt < P.n_chunks_total && nxt_free < NB) fu_wait_ge(P.consumed + (nxt_gi - P.ring), NB);
                claim[0] = nxt;
                claim[1] = nxt_gi;
                FU_T(tp5)
                FU_ACC(0, tp0, tp1) FU_ACC(1, tp1, tp2) FU_ACC(3, tp2, tp4) FU_ACC(4, tp4, tp5)
                FU_ACC(7, 0, 1)
            }
        }
    } else {
        // ------------------------------------------------------------------ B: build tiles
#ifdef MSBT_FU_SETMAXNREG
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(MSBT_FU_B_REGS));
#endif
        // Software-pipelined by one item: the claim of item i+1 and the L2 loads of its positions are issued
        // while item i's last tile drains, so that the only exposed latency per tile is the smem read-out of
        // the previous bulk store.
        const uint32_t tid = role_tid;
        const uint32_t saddr = (uint32_t)__cvta_generic_to_shared(tile);
        const uint32_t n_items = P.n_genomes * NB;
        const uint32_t tile_words32 = 1u << (P.tile_log2 - 5);
        const uint32_t tile_mask = (1u << P.tile_log2) - 1u;
        const uint32_t tiles_here = 1u << P.tiles_per_bucket_log2;
        bool store_pending = false;
        uint64_t evict_first;
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(evict_first));
        uint4 e[FU_REG_QUADS];  // every ring reservation is a whole quad, so validity is per 16-byte group
        uint32_t n_raw = 0;

#define FU_B_WANT(it) (((P.genomes[(it) / NB].n_bases + FU_CHUNK_BASES - 1) / FU_CHUNK_BASES) * (FU_P_THREADS / 32))
#define FU_B_WANT_G(g) (((P.genomes[(g)].n_bases + FU_CHUNK_BASES - 1) / FU_CHUNK_BASES) * (FU_P_THREADS / 32))
#define FU_B_ISSUE_LOADS(it)                                                                       \
    {                                                                                              \
        const uint32_t gi_ = claim[5], b_ = (it) - gi_ * NB, slot_ = claim[6];                     \
        const uint4 *src_ = (const uint4 *)(P.gbuf + ((uint64_t)slot_ * NB + b_) * P.gcap);        \
        n_raw = __ldcg(P.gcnt + (uint64_t)gi_ * NB + b_);                                          \
        _Pragma("unroll") for (int u = 0; u < FU_REG_QUADS; u++)                                   \
            e[u] = __ldcg(src_ + min(u * FU_B_THREADS + tid, (P.gcap >> 2) - 1u));                 \
    }
// one position: OR its bit into the tile if it belongs to tile `tf`, else OR 0 into the same word (branch-free)
#define FU_B_OR(x)                                                                                 \
    {                                                                                              \
        const uint32_t v_ = (((x) >> P.tile_log2) == tf) ? (1u << ((x) & 31)) : 0u;                \
        asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(saddr + (((x) >> 3) & word_mask)), "r"(v_) : "memory"); \
    }

        if (P.fast_b) {
            // ---- lean loop for the common geometry (bucket == one full 64 KiB tile): 3 barriers per tile and as few
            // instructions as possible -- next to hashing warps a B warp only issues once every ~30 cycles, so the
            // length of its instruction stream IS its latency.
            //   A  zero the tile                                   (tid 32 resolves the next claim meanwhile)   bar
            //   B  OR the positions in, fence.proxy.async                                                       bar
            //   C  tid 0: bulk store, then wait for its smem read-out;  all: prefetch the next item's positions bar
            constexpr uint32_t TILE_VEC = FU_TILE_BYTES / 16;
            const uint32_t word_mask = ((FU_TILE_BYTES * 8 - 1) >> 3) & ~3u;
            uint32_t pend = 0;
            if (tid == 32) {
                const uint32_t it = atomicAdd(P.sync + 1, 1u);
                pend = atomicAdd(P.sync + 1, 1u);
                const uint32_t g = __umulhi(it, P.nb_magic);
                if (it < n_items) fu_wait_ge(P.done + g, FU_B_WANT_G(g));
                claim[4] = it;
                claim[5] = g;
            }
            fu_bar(2, FU_B_THREADS);
            uint32_t item = claim[4], gi = claim[5];
#define FU_B_FAST_LOADS()                                                                              \
    {                                                                                                  \
        const uint32_t b_ = item - gi * NB;                                                            \
        const uint4 *src_ = (const uint4 *)(P.gbuf + ((uint64_t)(gi % P.ring) * NB + b_) * P.gcap);    \
        n_raw = __ldcg(P.gcnt + (uint64_t)gi * NB + b_);                                               \
        _Pragma("unroll") for (int u = 0; u < FU_REG_QUADS; u++)                                       \
            e[u] = __ldcg(src_ + min(u * FU_B_THREADS + tid, (P.gcap >> 2) - 1u));                     \
    }
#define FU_B_OR1(x) asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(saddr + (((x) >> 3) & word_mask)), "r"(1u << ((x) & 31)) : "memory");
            if (item < n_items) FU_B_FAST_LOADS()
            const uint32_t tpb_log2 = P.tiles_per_bucket_log2, tpb = 1u << tpb_log2;
            while (item < n_items) {
                uint32_t nxt = 0, nxt_gi = 0, nxt_done = 0;
                if (tid == 32) {
                    nxt = pend;
                    pend = atomicAdd(P.sync + 1, 1u);
                    nxt_gi = __umulhi(nxt, P.nb_magic);
                    if (nxt < n_items) nxt_done = fu_ld_acquire(P.done + nxt_gi);
                }
                const uint32_t cur_gi = gi;
                const uint32_t b = item - gi * NB;
                uint64_t *dst = P.filters + P.genomes[gi].filter_word_off + ((uint64_t)b << tpb_log2) * (FU_TILE_BYTES / 8);
                uint32_t n = 0;
                for (uint32_t sub = 0; sub < tpb; sub++) {
                    const bool last = sub + 1 == tpb;
                    FU_T(la0)
                    // A
                    uint4 *tv = (uint4 *)tile;
#pragma unroll
                    for (uint32_t i = 0; i < (TILE_VEC + FU_B_THREADS - 1) / FU_B_THREADS; i++)
                        if (i * FU_B_THREADS + tid < TILE_VEC) tv[i * FU_B_THREADS + tid] = make_uint4(0, 0, 0, 0);
                    if (sub == 0) n = min(n_raw, P.gcap);  // the prefetched loads land here
                    if (sub == 0 && tid == 32) {
                        if (nxt < n_items && nxt_done < FU_B_WANT_G(nxt_gi)) fu_wait_ge(P.done + nxt_gi, FU_B_WANT_G(nxt_gi));
                        claim[4] = nxt;
                        claim[5] = nxt_gi;
                    }
                    FU_T(la1)
                    fu_bar(2, FU_B_THREADS);
                    FU_T(la2)
                    // B
                    if (tpb == 1) {
#pragma unroll
                        for (int u = 0; u < FU_REG_QUADS; u++)
                            if ((u * FU_B_THREADS + tid) * 4u < n) { FU_B_OR1(e[u].x) FU_B_OR1(e[u].y) FU_B_OR1(e[u].z) FU_B_OR1(e[u].w) }
                    } else {
                        const uint32_t tf = (b << tpb_log2) + sub;  // a position of another tile ORs 0 into this one
#define FU_B_OR2(x)                                                                                           \
    {                                                                                                         \
        const uint32_t v_ = (((x) >> FU_TILE_LOG2) == tf) ? (1u << ((x) & 31)) : 0u;                          \
        asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(saddr + (((x) >> 3) & word_mask)), "r"(v_) : "memory"); \
    }
#pragma unroll
                        for (int u = 0; u < FU_REG_QUADS; u++)
                            if ((u * FU_B_THREADS + tid) * 4u < n) { FU_B_OR2(e[u].x) FU_B_OR2(e[u].y) FU_B_OR2(e[u].z) FU_B_OR2(e[u].w) }
                    }
                    if (n > FU_REG_ENTRIES * FU_B_THREADS) {  // oversized bucket (skewed input)
                        const uint32_t *src = P.gbuf + ((uint64_t)(cur_gi % P.ring) * NB + b) * P.gcap;
                        const uint32_t tf = (b << tpb_log2) + sub;
                        for (uint32_t i = FU_REG_ENTRIES * FU_B_THREADS + tid; i < n; i += FU_B_THREADS) {
                            const uint32_t p = __ldcg(src + i);
                            if ((p >> FU_TILE_LOG2) == tf) FU_B_OR1(p)
                        }
                    }
                    FU_T(la3)
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    fu_bar(2, FU_B_THREADS);
                    FU_T(la4)
                    // C
                    if (last) {
                        item = claim[4];
                        gi = claim[5];
                        if (item < n_items) FU_B_FAST_LOADS()
                    }
                    FU_T(la5)
                    if (tid == 0) {
                        asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(dst + (uint64_t)sub * (FU_TILE_BYTES / 8)), "r"(saddr), "r"(FU_TILE_BYTES), "l"(evict_first) : "memory");
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    }
                    FU_T(la6)
                    if (last && tid == 64) atomicAdd(P.consumed + cur_gi, 1u);
                    fu_bar(2, FU_B_THREADS);
                    FU_T(la7)
                    // lean loop phases: tid 0 (issues the store and waits for the read-out) and tid 96 (a plain worker lane)
                    if (tid == 0) { FU_ACC(8, la5, la6) FU_ACC(10, la0, la1) FU_ACC(11, la2, la3) FU_ACC(14, la4, la5) FU_ACC(15, 0, 1) FU_ACC(16, la0, la7) }
                    if (tid == 96) { FU_ACC(9, la1, la2) FU_ACC(12, la3, la4) FU_ACC(13, la6, la7) FU_ACC(17, la0, la7) }
                }
            }
#undef FU_B_OR2
#undef FU_B_OR1
#undef FU_B_FAST_LOADS
            return;
        }

        // Three lanes of different warps carry the long-latency control operations so that none of them
        // serialises behind another: tid 0 issues / drains the bulk stores, tid 32 claims items (two ahead, so
        // the atomic's round trip is never waited for) and checks `done`, tid 64 hands ring buckets back.
        uint32_t pend = 0;  // tid 32: the item after next
        if (tid == 32) {
            const uint32_t it = atomicAdd(P.sync + 1, 1u);
            pend = atomicAdd(P.sync + 1, 1u);
            if (it < n_items) fu_wait_ge(P.done + it / NB, FU_B_WANT(it));
            claim[4] = it;
            claim[5] = it / NB;
            claim[6] = (it / NB) % P.ring;
        }
        fu_bar(2, FU_B_THREADS);
        uint32_t item = claim[4];
        if (item < n_items) FU_B_ISSUE_LOADS(item)

        while (item < n_items) {
            FU_T(tbA)
            uint32_t nxt = 0, nxt_gi = 0, nxt_done = 0;
            if (tid == 32) {
                nxt = pend;
                pend = atomicAdd(P.sync + 1, 1u);                               // consumed next iteration
                nxt_gi = nxt / NB;
                if (nxt < n_items) nxt_done = fu_ld_acquire(P.done + nxt_gi);    // consumed after the last tile
            }
            const uint32_t gi = claim[5], b = item - gi * NB;
            const uint32_t slot = claim[6];
            const uint64_t fw = P.genomes[gi].filter_word_off;
            const uint32_t *src = P.gbuf + ((uint64_t)slot * NB + b) * P.gcap;
            uint32_t n = 0;
            FU_T(tbB)
            if (tid == 0) { FU_ACC(17, tbA, tbB) }
            if (tid == 32) { FU_ACC(18, tbA, tbB) }
            for (uint32_t sub = 0; sub < tiles_here; sub++) {
                const uint32_t tf = (b << P.tiles_per_bucket_log2) + sub;   // tile index inside the filter
                const uint64_t w0 = ((uint64_t)tf << P.tile_log2) >> 6;     // its first u64 word
                if (w0 >= P.words) break;
                FU_T(tb0)
                if (store_pending && tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                FU_T(tb1)
                fu_bar(2, FU_B_THREADS);
                FU_T(tb2)
                uint4 *tv = (uint4 *)tile;
                for (uint32_t i = tid; i < (tile_words32 + 3) / 4; i += FU_B_THREADS) tv[i] = make_uint4(0, 0, 0, 0);
                if (sub == 0) {
                    n = min(n_raw, P.gcap);  // the prefetched loads land here
                }
                fu_bar(2, FU_B_THREADS);
                FU_T(tb3)
                if (n) {
                    const uint32_t word_mask = (tile_mask >> 3) & ~3u;
                    // groups past the end of the bucket are skipped (a thread-uniform branch per 16-byte group)
                    if (tiles_here == 1) {  // bucket == tile: every position belongs here, no test at all
#define FU_B_OR1(x) asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(saddr + (((x) >> 3) & word_mask)), "r"(1u << ((x) & 31)) : "memory");
#pragma unroll
                        for (int u = 0; u < FU_REG_QUADS; u++)
                            if ((u * FU_B_THREADS + tid) * 4u < n) { FU_B_OR1(e[u].x) FU_B_OR1(e[u].y) FU_B_OR1(e[u].z) FU_B_OR1(e[u].w) }
                        for (uint32_t i = FU_REG_ENTRIES * FU_B_THREADS + tid; i < n; i += FU_B_THREADS) {  // oversized buckets
                            const uint32_t p = __ldcg(src + i);
                            FU_B_OR1(p)
                        }
#undef FU_B_OR1
                    } else {
#pragma unroll
                        for (int u = 0; u < FU_REG_QUADS; u++)
                            if ((u * FU_B_THREADS + tid) * 4u < n) { FU_B_OR(e[u].x) FU_B_OR(e[u].y) FU_B_OR(e[u].z) FU_B_OR(e[u].w) }
                        for (uint32_t i = FU_REG_ENTRIES * FU_B_THREADS + tid; i < n; i += FU_B_THREADS) {  // oversized buckets
                            const uint32_t p = __ldcg(src + i);
                            FU_B_OR(p)
                        }
                    }
                }
                FU_T(tb4)
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                fu_bar(2, FU_B_THREADS);
                FU_T(tb5)
                if (tid == 0) { FU_ACC(8, tb0, tb1) FU_ACC(9, tb1, tb2) FU_ACC(10, tb2, tb3) FU_ACC(11, tb3, tb4) FU_ACC(12, tb4, tb5) FU_ACC(15, 0, 1) }
                const uint64_t nw = min((uint64_t)(tile_words32 / 2), P.words - w0);
                uint64_t *dst = P.filters + fw + w0;
                const uint32_t bytes = (uint32_t)(nw * 8);
                if (((bytes & 15) == 0) && ((((uintptr_t)dst) & 15) == 0)) {
                    if (tid == 0) {
                        // evict_first: the finished tile streams to HBM and must not push the ring out of L2
                        asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(dst), "r"(saddr), "r"(bytes), "l"(evict_first) : "memory");
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                    store_pending = true;
                } else {
                    const uint64_t *tw = (const uint64_t *)tile;
                    for (uint64_t i = tid; i < nw; i += FU_B_THREADS) dst[i] = tw[i];
                }
                FU_T(tb5b)
                if (tid == 0) { FU_ACC(16, tb5, tb5b) }
            }
            // Every position of this bucket has been consumed (the barrier before the last store).
            FU_T(tb6)
            if (tid == 32) {
                if (nxt < n_items && nxt_done < FU_B_WANT(nxt)) fu_wait_ge(P.done + nxt_gi, FU_B_WANT(nxt));
                claim[4] = nxt;
                claim[5] = nxt_gi;
                claim[6] = nxt_gi % P.ring;
            }
            fu_bar(2, FU_B_THREADS);
            FU_T(tb7)
            item = claim[4];
            if (item < n_items) FU_B_ISSUE_LOADS(item)
            FU_T(tb8)
            if (tid == 0) { FU_ACC(13, tb6, tb7) FU_ACC(14, tb7, tb8) }
            if (tid == 64) atomicAdd(P.consumed + gi, 1u);  // hand the ring bucket back (nothing to order: B wrote nothing P reads)
        }
#undef FU_B_OR
#undef FU_B_ISSUE_LOADS
#undef FU_B_WANT_G
#undef FU_B_WANT
        if (store_pending && tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}
