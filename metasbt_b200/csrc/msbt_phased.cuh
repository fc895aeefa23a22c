// msbt_phased.cuh -- K1 variant 4 ("phased"): one persistent kernel per batch, 1 CTA per SM, 1024 threads that
// all alternate between two phases over the SAME shared-memory area (replaces `howdesbt makebf`,
// /root/reference/metasbt/core.py:3920-3931, for every genome of the batch):
//
//   HASH  phase  claim a chunk of PH_CHUNK_BASES k-mer starts of one genome, hash every canonical k-mer (32 per
//                thread), stage its position by bucket in shared memory (bucket = pos >> bucket_shift, at most
//                2048 buckets per filter, one staging row per bucket), then flush every row with one reservation
//                and aligned 16-byte stores into an L2-resident scratch ring (one slot per genome in flight).
//                Issue-bound: 64-bit MurmurHash3 on a 32-bit datapath, ~90 instructions per k-mer.
//   BUILD phase  claim a batch of (genome, bucket) items of genomes whose chunks are all flushed; for each tile
//                (<= 64 KiB) of the bucket: the tile is zeroed in shared memory, the bucket's positions are OR-ed
//                in with shared-memory REDs, and the finished tile goes to HBM with ONE TMA bulk store
//                (cp.async.bulk.global.shared::cta).  Three tile buffers rotate, so the bulk store of tile t
//                drains while tiles t+1, t+2 are zeroed and filled; the stores of a phase keep draining during
//                the next HASH phase's first instructions.
//
// The filter is written exactly once (zeros included): HBM traffic per genome = packed input + filter.
// Why phases instead of concurrent roles (variant 3): the tile-build chain is latency-bound and starves next
// to hashing warps (measured CPI ~30), while here all 32 warps shorten every link of that chain and no warp
// ever waits for the other role.  The staging area is dead between a flush and the next chunk, which is
// exactly when the tiles need it.
//
// Flow control through global counters zeroed by one memset before the launch:
//   sync[0] next chunk, sync[32] next item (separate 128-byte lines)
//   done[g]      warps that flushed a chunk of genome g   (items of g are ready at n_chunks(g) * 32)
//   consumed[g]  items of genome g built                  (ring slot of g reusable at NB)
//   dirty[g]     a ring bucket of g overflowed (skewed input): g is redone by the direct kernel afterwards
// Claims are handed out in order; a CTA prefers BUILD whenever its item is ready, else HASH if the ring slot of
// its chunk is free, else polls.  Lowest unfinished genome g*: its chunks are never ring-blocked and its items
// are ready as soon as its chunks are done, so some CTA always progresses.
#pragma once

constexpr int PH_THREADS = 1024;
constexpr uint32_t PH_CHUNK_BASES = PH_THREADS * 32;
constexpr uint32_t PH_MAX_BUCKETS = 2048;
constexpr uint32_t PH_TILE_LOG2 = 19;  // <= 2^19 bits = 64 KiB per tile
constexpr uint32_t PH_TILE_BYTES = 1u << (PH_TILE_LOG2 - 3);
constexpr uint32_t PH_NBUF = 3;        // rotating tile buffers inside the area
// dynamic shared memory: [cnt 2048][area][ctrl 16]; the area holds (NB + 1) staging rows or 3 tile buffers
constexpr uint32_t PH_AREA_WORDS = 54528;  // 218,112 B  (>= 3 * 64 KiB; 26 words per row at 2049 rows)
constexpr size_t PH_SMEM_BYTES = (size_t)PH_MAX_BUCKETS * 4 + (size_t)PH_AREA_WORDS * 4 + 64;
constexpr uint32_t PH_ITEM_BATCH = 16;  // items claimed per BUILD phase
constexpr int PH_REG_QUADS = 2;         // 16-byte groups of bucket positions per thread kept in registers
constexpr uint32_t PH_NONE = 0xFFFFFFFFu;
static_assert(PH_AREA_WORDS * 4 >= PH_NBUF * PH_TILE_BYTES, "area must hold the tile buffers");
static_assert(PH_SMEM_BYTES <= 232448, "227 KiB of shared memory per CTA on sm_100");

struct PhasedParams {
    const uint64_t *packed;
    const uint32_t *run_ends;
    const DevGenome *genomes;
    uint32_t n_genomes;
    uint32_t n_chunks_total;
    uint32_t n_items;        // n_genomes * NB
    uint32_t bucket_shift;   // bucket = pos >> bucket_shift
    uint32_t NB;             // buckets per filter (<= PH_MAX_BUCKETS)
    uint32_t cap;            // staging entries per bucket per chunk (multiple of 4)
    uint32_t quads;          // cap / 4
    uint32_t tile_log2;      // min(bucket_shift, PH_TILE_LOG2)
    uint32_t tiles_per_bucket_log2;
    uint32_t gcap;           // ring entries per bucket (multiple of 8)
    uint32_t ring;           // genomes in flight
    uint64_t words;          // u64 words per filter
    uint32_t *sync;          // [0] next chunk, [32] next item
    uint32_t *done;          // [n_genomes]
    uint32_t *consumed;      // [n_genomes]
    uint32_t *dirty;         // [n_genomes]
    uint32_t *gcnt;          // [n_genomes][NB] entries appended per (genome, bucket)
    uint32_t *gbuf;          // [ring][NB][gcap]
    uint64_t *filters;
};

__device__ __forceinline__ uint32_t ph_ld_relaxed(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t ph_chunks_of(const DevGenome &g) { return (g.n_bases + PH_CHUNK_BASES - 1) / PH_CHUNK_BASES; }

#define PH_HASH(J)                                                                                                   \
    {                                                                                                                \
        uint64_t lo_, hi_;                                                                                           \
        kb.template get<(H * 4 + J)>(lo_, hi_);                                                                      \
        const uint64_t p_ = msbt_position_v<POW2>(msbt_kmer_hash_v<HV>(lo_, hi_, len, spec.pos.seed), spec.pos);     \
        pos4[J] = (uint32_t)p_;                                                                                      \
        ok4[J] = (AV || ((vg >> (H * 4 + J)) & 1u)) && (FULL || p_ < spec.pos.num_bits);                             \
        b4[J] = ok4[J] ? (pos4[J] >> P.bucket_shift) : P.NB;                                                         \
    }
#define PH_RANK(J) r4[J] = (AV && FULL) ? atoms_inc(cnt_s + b4[J] * 4u) : atoms_inc_if(cnt_s + b4[J] * 4u, ok4[J]);
#define PH_STORE(J)                                                                   \
    sts_if(stage_s + (b4[J] * P.cap + r4[J]) * 4u, pos4[J], r4[J] < P.cap);           \
    over |= (r4[J] >= P.cap);
#define PH_COLD(J) \
    if (r4[J] >= P.cap) fu_append_cold(gc, gb, P.gcap, b4[J], pos4[J], dirty_g);

// AV: all 32 k-mer starts of the thread's word are valid (no run boundary inside); with FULL (modulus <= num_bits)
// the k-mer then needs no predicate at all.
template <int KW, int HV, bool POW2, bool FULL, bool AV, int H>
__device__ __forceinline__ void ph_half_step(const KmerBlock<KW> &kb, uint32_t vg, const KmerSpec &spec, const PhasedParams &P,
                                             uint32_t len, uint32_t cnt_s, uint32_t stage_s, uint32_t *gc, uint32_t *gb,
                                             uint32_t *dirty_g) {
    uint32_t pos4[4], b4[4], r4[4];
    bool ok4[4];
    PH_HASH(0) PH_HASH(1) PH_HASH(2) PH_HASH(3)
    PH_RANK(0) PH_RANK(1) PH_RANK(2) PH_RANK(3)
    bool over = false;
    PH_STORE(0) PH_STORE(1) PH_STORE(2) PH_STORE(3)
    if (over) { PH_COLD(0) PH_COLD(1) PH_COLD(2) PH_COLD(3) }  // staging row full: append straight to the ring bucket
}

template <int KW, int HV, bool POW2, bool FULL>
__global__ void __launch_bounds__(PH_THREADS, 1)
kmer_bloom_phased_kernel(PhasedParams P, KmerSpec spec) {
    extern __shared__ __align__(128) uint32_t ph_smem[];
    uint32_t *area = ph_smem;                           // 128-byte aligned: staging rows / tile buffers
    uint32_t *cnt = ph_smem + PH_AREA_WORDS;            // [NB] entries staged per bucket
    volatile uint32_t *ctrl = cnt + PH_MAX_BUCKETS;     // decisions of thread 0
    const uint32_t tid = threadIdx.x;
    const uint32_t NB = P.NB;
    const uint32_t cnt_s = (uint32_t)__cvta_generic_to_shared(cnt);
    const uint32_t area_s = (uint32_t)__cvta_generic_to_shared(area);
    const uint32_t k = spec.k, len = (k + 3) >> 2;
    enum { ACT_EXIT = 0, ACT_HASH = 1, ACT_BUILD = 2 };

    // ---- coordinator state lives in shared memory (only thread 0 touches it; keeps it out of everyone's registers):
    //   ctrl[4] claimed chunk, [5] its genome, [6] the chunk after it (claim in flight)
    //   ctrl[7], [8] claimed item batch [ib, ie), [9] the batch after it
    if (tid == 0) {
        const uint32_t ci = atomicAdd(P.sync + 0, 1u);
        ctrl[6] = atomicAdd(P.sync + 0, 1u);
        const uint32_t ib = atomicAdd(P.sync + 32, PH_ITEM_BATCH);
        ctrl[9] = atomicAdd(P.sync + 32, PH_ITEM_BATCH);
        uint32_t ci_g = 0;
        if (ci < P.n_chunks_total) {
            uint32_t lo = 0, hi = P.n_genomes;  // last genome with fchunk_begin <= ci
            while (hi - lo > 1) {
                const uint32_t mid = (lo + hi) >> 1;
                if (P.genomes[mid].fchunk_begin <= ci) lo = mid; else hi = mid;
            }
            ci_g = lo;
        }
        ctrl[4] = ci; ctrl[5] = ci_g; ctrl[7] = ib; ctrl[8] = min(ib + PH_ITEM_BATCH, P.n_items);
    }
    for (uint32_t b = tid; b < NB; b += PH_THREADS) cnt[b] = 0;
    uint32_t buf_next = 0;        // next tile buffer (uniform across the CTA)
    bool stores_in_flight = false;

    while (true) {
        FU_T(td0)
        // ------------------------------------------------------------------ decide (thread 0)
        if (tid == 0) {
            uint32_t act = ACT_EXIT, a0 = 0, a1 = 0, a2 = 0;
            uint32_t ci = ctrl[4], ci_g = ctrl[5], ib = ctrl[7], ie = ctrl[8];
            while (true) {
                bool progress = false;
                if (ib < P.n_items) {
                    const uint32_t g0 = ib / NB;
                    if (ph_ld_relaxed(P.done + g0) >= ph_chunks_of(P.genomes[g0]) * (PH_THREADS / 32)) {
                        // ready prefix of the batch: up to the end of genome g0, further if the next genome is done too
                        uint32_t end = min(ie, (g0 + 1) * NB);
                        (void)fu_ld_acquire(P.done + g0);  // order the ring reads of this phase after the flushes
                        if (end < ie && ph_ld_relaxed(P.done + g0 + 1) >= ph_chunks_of(P.genomes[g0 + 1]) * (PH_THREADS / 32)) {
                            end = ie;
                            (void)fu_ld_acquire(P.done + g0 + 1);
                        }
                        act = ACT_BUILD; a0 = ib; a1 = end; a2 = g0;
                        ib = end;
                        if (ib == ie) {  // batch finished after this phase: move to the pending one
                            ib = ctrl[9];
                            ie = min(ib + PH_ITEM_BATCH, P.n_items);
                            ctrl[9] = atomicAdd(P.sync + 32, PH_ITEM_BATCH);
                        }
                        progress = true;
                    }
                }
                if (!progress && ci < P.n_chunks_total) {
                    if (ci_g < P.ring || ph_ld_relaxed(P.consumed + (ci_g - P.ring)) >= NB) {
                        act = ACT_HASH; a0 = ci; a1 = ci_g;
                        ci = ctrl[6];
                        ctrl[6] = atomicAdd(P.sync + 0, 1u);
                        if (ci < P.n_chunks_total)  // claims are monotonic: scan forward from the current genome
                            while (ci_g + 1 < P.n_genomes && P.genomes[ci_g + 1].fchunk_begin <= ci) ci_g++;
                        progress = true;
                    }
                }
                if (progress) break;
                if (ib >= P.n_items && ci >= P.n_chunks_total) { act = ACT_EXIT; break; }
#ifdef MSBT_FU_TRACE
                if (blockIdx.x == 0) {
                    const bool ring_blocked = ci < P.n_chunks_total, item_blocked = ib < P.n_items;
                    atomicAdd(&fu_trace[16 + (ring_blocked ? 1 : 0) + (item_blocked ? 2 : 0)], 1ull);
                }
#endif
                __nanosleep(500);
            }
            ctrl[4] = ci; ctrl[5] = ci_g; ctrl[7] = ib; ctrl[8] = ie;
            ctrl[0] = act; ctrl[1] = a0; ctrl[2] = a1; ctrl[3] = a2;
        }
        __syncthreads();  // decision visible; staging counters zero; previous phase complete
        const uint32_t act = ctrl[0], a0 = ctrl[1], a1 = ctrl[2], a2 = ctrl[3];
        if (act == ACT_EXIT) break;
        FU_T(td1)
        if (tid == 0) { FU_ACC(0, td0, td1) }

        if (act == ACT_HASH) {
            // ================================================================== HASH phase: chunk a0 of genome a1
            const uint32_t gi = a1;
            const DevGenome g = P.genomes[gi];
            const uint32_t slot = gi % P.ring;
            uint32_t *gc = P.gcnt + (uint64_t)gi * NB;
            uint32_t *gb = P.gbuf + (uint64_t)slot * NB * P.gcap;
            uint32_t *dirty_g = P.dirty + gi;
            if (stores_in_flight) {  // the staging rows overwrite the tile buffers: their bulk stores must have read them
                if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                __syncthreads();
                stores_in_flight = false;
            }
            FU_T(th1)
            const uint32_t w = (a0 - g.fchunk_begin) * PH_THREADS + tid;
            const uint32_t s0 = w * 32;
            uint32_t vm = 0;
            if (s0 < g.n_bases) vm = valid_mask(P.run_ends + g.run_off, g.n_runs, s0, k);
            if (vm == 0xFFFFFFFFu) {
                KmerBlock<KW> kb;
                kb.init(P.packed + g.packed_word_off + w, k);
#pragma unroll 1
                for (uint32_t grp = 0; grp < 4; grp++) {
                    ph_half_step<KW, HV, POW2, FULL, true, 0>(kb, 0xFFu, spec, P, len, cnt_s, area_s, gc, gb, dirty_g);
                    ph_half_step<KW, HV, POW2, FULL, true, 1>(kb, 0xFFu, spec, P, len, cnt_s, area_s, gc, gb, dirty_g);
                    kb.advance8();
                }
            } else if (vm) {
                KmerBlock<KW> kb;
                kb.init(P.packed + g.packed_word_off + w, k);
#pragma unroll 1
                for (uint32_t grp = 0; grp < 4; grp++) {
                    const uint32_t vg = (vm >> (8 * grp)) & 0xFFu;
                    ph_half_step<KW, HV, POW2, FULL, false, 0>(kb, vg, spec, P, len, cnt_s, area_s, gc, gb, dirty_g);
                    ph_half_step<KW, HV, POW2, FULL, false, 1>(kb, vg, spec, P, len, cnt_s, area_s, gc, gb, dirty_g);
                    kb.advance8();
                }
            }
            __syncthreads();  // the chunk is staged
            FU_T(th2)
            // Flush: a thread owns whole staging rows.  One ring reservation per (chunk, bucket), rounded up to whole
            // quads; the reservations of all its rows are issued before the first is consumed.  Rows are copied with
            // aligned 16-byte loads/stores; pad slots repeat the row's first entry (OR is idempotent).  The counter is
            // zeroed for the next chunk.
            {
                constexpr int ROWS = PH_MAX_BUCKETS / PH_THREADS;
                uint32_t nrow[ROWS], orow[ROWS];
#pragma unroll
                for (int j = 0; j < ROWS; j++) {
                    const uint32_t b = tid + j * PH_THREADS;
                    uint32_t n = 0;
                    if (b < NB) { n = min(cnt[b], P.cap); cnt[b] = 0; }
                    nrow[j] = n;
                    orow[j] = n ? atomicAdd(gc + b, (n + 3u) & ~3u) : 0u;
                }
                FU_T(tf1)
                if (tid == 0) { FU_ACC(5, th2, tf1) }
                const uint4 *stage4 = (const uint4 *)area;
                uint4 *gb4 = (uint4 *)gb;
#pragma unroll
                for (int j = 0; j < ROWS; j++) {
                    const uint32_t n = nrow[j];
                    if (n) {
                        const uint32_t b = tid + j * PH_THREADS;
                        const uint32_t nq = (n + 3u) >> 2;
                        if (orow[j] + nq * 4u <= P.gcap) {
                            const uint4 *src4 = stage4 + b * P.quads;
                            uint4 *dst4 = gb4 + (((uint64_t)b * P.gcap + orow[j]) >> 2);
                            const uint32_t first = src4[0].x;
                            for (uint32_t q = 0; q + 1 < nq; q++) dst4[q] = src4[q];
                            uint4 v = src4[nq - 1];
                            const uint32_t have = n - (nq - 1) * 4u;
                            if (have < 2) v.y = first;
                            if (have < 3) v.z = first;
                            if (have < 4) v.w = first;
                            dst4[nq - 1] = v;
                        } else {
                            *dirty_g = 1u;  // ring bucket full (skewed input): the direct kernel redoes this genome
                            // pad the tail this reservation leaves unwritten (the BUILD phase reads min(count, gcap) entries)
                            const uint32_t first = stage4[b * P.quads].x;
                            for (uint32_t o = orow[j]; o + 4u <= P.gcap; o += 4u)
                                *(uint4 *)(gb + (uint64_t)b * P.gcap + o) = make_uint4(first, first, first, first);
                        }
                    }
                }
            }
            // publish per warp: done[g] counts warps, so no CTA barrier sits between the flush and the publish
            FU_T(tf2)
            __threadfence();
            __syncwarp();
            if ((tid & 31u) == 0) atomicAdd(P.done + gi, 1u);
            FU_T(th3)
            if (tid == 0) { FU_ACC(1, td1, th1) FU_ACC(2, th1, th2) FU_ACC(3, th2, th3) FU_ACC(4, 0, 1) FU_ACC(6, th2, tf2) }
        } else {
            // ================================================================== BUILD phase: items [a0, a1), first genome a2
            uint32_t gi = a2, b = a0 - a2 * NB;
            const uint32_t tile_words32 = 1u << (P.tile_log2 - 5);
            const uint32_t tile_mask = (1u << P.tile_log2) - 1u;
            const uint32_t word_mask = (tile_mask >> 3) & ~3u;
            const uint32_t tiles_here = 1u << P.tiles_per_bucket_log2;
            const uint32_t tile_vec = (tile_words32 + 3) / 4;  // uint4 per tile
            uint4 e[PH_REG_QUADS];
            uint32_t n_raw;
#define PH_ISSUE_LOADS(gi_, b_)                                                                        \
    {                                                                                                  \
        const uint4 *src_ = (const uint4 *)(P.gbuf + ((uint64_t)((gi_) % P.ring) * NB + (b_)) * P.gcap); \
        n_raw = __ldcg(P.gcnt + (uint64_t)(gi_) * NB + (b_));                                          \
        _Pragma("unroll") for (int u = 0; u < PH_REG_QUADS; u++)                                       \
            e[u] = __ldcg(src_ + min(u * PH_THREADS + tid, (P.gcap >> 2) - 1u));                       \
    }
#define PH_OR(x, buf_s)                                                                                \
    {                                                                                                  \
        const uint32_t v_ = (MULTI && (((x) >> P.tile_log2) != tf)) ? 0u : (1u << ((x) & 31));         \
        asm volatile("red.shared.or.b32 [%0], %1;" ::"r"((buf_s) + (((x) >> 3) & word_mask)), "r"(v_) : "memory"); \
    }
            PH_ISSUE_LOADS(gi, b)
            // zero the first tile buffer of the phase (the staging rows are dead after the flush)
            {
                uint4 *tv = (uint4 *)(area + buf_next * (PH_TILE_BYTES / 4));
                for (uint32_t i = tid; i < tile_vec; i += PH_THREADS) tv[i] = make_uint4(0, 0, 0, 0);
            }
            FU_T(tb0)
            if (tid == 0) { FU_ACC(8, td1, tb0) FU_ACC(13, 0, 1) }
            for (uint32_t item = a0; item < a1; item++) {
                const uint32_t n = min(n_raw, P.gcap);
                // groups past the end of the bucket are skipped below (never OR a repeated position: REDs of one warp
                // instruction to the same word serialise)
                const bool v0 = tid * 4u < n, v1 = (PH_THREADS + tid) * 4u < n;
                const uint4 e0 = e[0], e1 = e[1];
                const uint32_t cur_gi = gi, cur_b = b;
                const uint32_t *src = P.gbuf + ((uint64_t)(cur_gi % P.ring) * NB + cur_b) * P.gcap;
                const uint64_t fw = P.genomes[cur_gi].filter_word_off;
                // next item of the batch (consecutive buckets; the genome advances when the bucket index wraps)
                b++;
                if (b == NB) { b = 0; gi++; }
                if (item + 1 < a1) PH_ISSUE_LOADS(gi, b)   // prefetch: lands while this item's tiles are built
                for (uint32_t sub = 0; sub < tiles_here; sub++) {
                    const uint32_t tf = (cur_b << P.tiles_per_bucket_log2) + sub;  // tile index inside the filter
                    const uint64_t w0 = ((uint64_t)tf << P.tile_log2) >> 6;        // its first u64 word
                    if (w0 >= P.words) break;
                    const uint32_t buf = buf_next;
                    const uint32_t nbuf = (buf + 1 == PH_NBUF) ? 0 : buf + 1;
                    const uint32_t buf_s = area_s + buf * PH_TILE_BYTES;
                    // the buffer zeroed below was last read by the store issued PH_NBUF - 1 tiles ago
                    FU_T(tt0)
                    if (stores_in_flight && tid == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                    FU_T(tt1)
                    __syncthreads();  // S1: this tile's buffer is zero; the next buffer is free
                    FU_T(tt2)
                    if (n) {
                        if (tiles_here > 1) {
                            constexpr bool MULTI = true;
                            if (v0) { PH_OR(e0.x, buf_s) PH_OR(e0.y, buf_s) PH_OR(e0.z, buf_s) PH_OR(e0.w, buf_s) }
                            if (v1) { PH_OR(e1.x, buf_s) PH_OR(e1.y, buf_s) PH_OR(e1.z, buf_s) PH_OR(e1.w, buf_s) }
                            for (uint32_t i = PH_REG_QUADS * 4 * PH_THREADS + tid; i < n; i += PH_THREADS) {  // oversized buckets
                                const uint32_t p = __ldcg(src + i);
                                PH_OR(p, buf_s)
                            }
                        } else {
                            constexpr bool MULTI = false;
                            const uint32_t tf = 0;
                            (void)tf;
                            if (v0) { PH_OR(e0.x, buf_s) PH_OR(e0.y, buf_s) PH_OR(e0.z, buf_s) PH_OR(e0.w, buf_s) }
                            if (v1) { PH_OR(e1.x, buf_s) PH_OR(e1.y, buf_s) PH_OR(e1.z, buf_s) PH_OR(e1.w, buf_s) }
                            for (uint32_t i = PH_REG_QUADS * 4 * PH_THREADS + tid; i < n; i += PH_THREADS) {
                                const uint32_t p = __ldcg(src + i);
                                PH_OR(p, buf_s)
                            }
                        }
                    }
                    {   // zero the next buffer while this one is being filled
                        uint4 *tv = (uint4 *)(area + nbuf * (PH_TILE_BYTES / 4));
                        for (uint32_t i = tid; i < tile_vec; i += PH_THREADS) tv[i] = make_uint4(0, 0, 0, 0);
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncthreads();  // S2: the tile is complete
                    FU_T(tt3)
                    const uint64_t nw = min((uint64_t)(tile_words32 / 2), P.words - w0);
                    uint64_t *dst = P.filters + fw + w0;
                    const uint32_t bytes = (uint32_t)(nw * 8);
                    if (((bytes & 15) == 0) && ((((uintptr_t)dst) & 15) == 0)) {
                        if (tid == 0) {
                            // evict_first: the finished tile streams to HBM and must not push the ring out of L2
                            uint64_t evict_first;
                            asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(evict_first));
                            asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(dst), "r"(buf_s), "r"(bytes), "l"(evict_first) : "memory");
                            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                        }
                        stores_in_flight = true;
                    } else {
                        const uint64_t *tw = (const uint64_t *)(area + buf * (PH_TILE_BYTES / 4));
                        for (uint64_t i = tid; i < nw; i += PH_THREADS) dst[i] = tw[i];
                        __syncthreads();  // the plain copy has read the buffer before it is zeroed again
                    }
                    buf_next = nbuf;
                    FU_T(tt4)
                    if (tid == 0) { FU_ACC(9, tt0, tt1) FU_ACC(10, tt1, tt2) FU_ACC(11, tt2, tt3) FU_ACC(12, tt3, tt4) FU_ACC(14, 0, 1) }
                }
                // every position of this bucket has been consumed (S2 of its last tile)
                if (tid == 0) atomicAdd(P.consumed + cur_gi, 1u);
            }
#undef PH_OR
#undef PH_ISSUE_LOADS
        }
    }
    if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
