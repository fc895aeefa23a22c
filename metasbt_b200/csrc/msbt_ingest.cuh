// msbt_ingest.cuh -- FASTA text -> 2-bit packed genomes + run tables ON THE DEVICE (the stage
// immediately in front of K1; replaces howdesbt's own FASTA reader behind `makebf <fasta>` / `query <fasta>`,
// /root/reference/metasbt/core.py:3928 and :2046).  Same semantics, byte for byte, as the host packer
// msbt_fasta_pack (msbt_kernels.cu), which the tests use as the checker:
//   * a line whose FIRST byte is '>' is a header: skipped entirely, and it ends the current run;
//   * A C G T (either case) are bases; ' ' '\t' '\r' are skipped; '\n' only marks a line start;
//   * every other byte ends the current run; a run shorter than `min_run` (= k) is dropped;
//   * packed stream = the bases of the kept runs only, 32 per u64, first base in the top two bits;
//     run_ends = cumulative end offsets of the kept runs inside that stream.
// A batch of files is one concatenated text; files are independent (ranks restart at every file).
//
// Five small passes, each one CTA per 4 KiB chunk of a file (256 threads x 16 bytes, text staged in shared memory):
//   A1  line summaries  -> does the chunk contain a newline, and in which line state does it end
//   (scan over chunks)  -> line state at the start of every chunk
//   A2  counts          -> bases and break events per chunk            (scan -> per-chunk ranks, per-file totals)
//   B   break table     -> Bt[j] = number of bases in front of the j-th break event of the file
//   C   run scan        -> per run: kept?, bases of kept runs in front of it; writes run_ends  (one CTA per file)
//   D   emit            -> every kept base ORs its two bits into the packed stream (<= 2 atomics per 16 bytes)
#pragma once

constexpr int FP_THREADS = 256;
constexpr uint32_t FP_BYTES_PER_THREAD = 16;
constexpr uint32_t FP_CHUNK = FP_THREADS * FP_BYTES_PER_THREAD;  // 4096 bytes
enum : uint32_t { FP_ST_SEQ = 0, FP_ST_HDR = 1, FP_ST_LINE_START = 2 };
// chunk summary codes of pass A1
enum : uint32_t { FP_PASS_SEQ = 0, FP_PASS_HDR = 1, FP_SET_SEQ = 2, FP_SET_HDR = 3, FP_SET_LINE_START = 4 };

struct FpFile {
    uint64_t text_off;     // first byte of the file in the concatenated text
    uint64_t text_len;
    uint32_t chunk_begin;  // first chunk of the file
    uint32_t n_chunks;
    uint64_t packed_word_off;
    uint64_t run_off;      // first entry of the file in run_ends
    uint64_t bt_off;       // first entry of the file in the break table (filled in after pass A2)
};

__device__ __forceinline__ uint32_t fp_find_file(const FpFile *__restrict__ files, uint32_t n, uint32_t chunk) {
    uint32_t lo = 0, hi = n;  // last file with chunk_begin <= chunk
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (files[mid].chunk_begin <= chunk) lo = mid; else hi = mid;
    }
    return lo;
}

__device__ __forceinline__ uint32_t fp_code(uint8_t c) {  // 0..3 base, 4 skip, 5 newline, 6 break
    switch (c) {
        case 'A': case 'a': return 0;
        case 'C': case 'c': return 1;
        case 'G': case 'g': return 2;
        case 'T': case 't': return 3;
        case ' ': case '\t': case '\r': return 4;
        case '\n': return 5;
        default: return 6;
    }
}

// Stage the chunk in shared memory and work out the line state at the first byte of every thread.
// Returns the thread's byte range [o, e) inside the chunk; `hdr` / `at_ls` = state at byte o.
struct FpThread {
    uint32_t o, e;
    bool hdr, at_ls;
};

__device__ __forceinline__ FpThread fp_stage(const uint8_t *__restrict__ text, const FpFile &f, uint32_t local_chunk,
                                             uint32_t state_in, uint8_t *sm_text, int *sm_scan, uint32_t &chunk_len) {
    const uint64_t base = f.text_off + (uint64_t)local_chunk * FP_CHUNK;
    chunk_len = (uint32_t)min((uint64_t)FP_CHUNK, f.text_len - (uint64_t)local_chunk * FP_CHUNK);
    const uint32_t t = threadIdx.x;
    FpThread th;
    th.o = min(t * FP_BYTES_PER_THREAD, chunk_len);
    th.e = min(th.o + FP_BYTES_PER_THREAD, chunk_len);
    for (uint32_t i = th.o; i < th.e; i++) sm_text[i] = text[base + i];
    // last newline inside my bytes (chunk index) or -1, then an inclusive max-scan across the CTA
    int last = -1;
    for (uint32_t i = th.o; i < th.e; i++)
        if (sm_text[i] == '\n') last = (int)i;
    int incl = last;
    const int lane = t & 31, wid = t >> 5;
    for (int d = 1; d < 32; d <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl = max(incl, v);
    }
    if (lane == 31) sm_scan[wid] = incl;
    __syncthreads();  // also publishes sm_text
    int before = -1;  // last newline in front of my first byte
    for (int w = 0; w < wid; w++) before = max(before, sm_scan[w]);
    const int prev_lane = __shfl_up_sync(0xffffffffu, incl, 1);
    if (lane > 0) before = max(before, prev_lane);
    int line_start = before >= 0 ? before + 1 : (state_in == FP_ST_LINE_START ? 0 : -1);
    th.hdr = false;
    th.at_ls = false;
    if (line_start < 0) th.hdr = (state_in == FP_ST_HDR);
    else if ((uint32_t)line_start == th.o) th.at_ls = true;
    else th.hdr = (sm_text[line_start] == '>');
    return th;
}

// ---- A1: line summary of every chunk
__global__ void __launch_bounds__(FP_THREADS)
fp_line_summary_kernel(const uint8_t *__restrict__ text, const FpFile *__restrict__ files, uint32_t n_files, uint32_t n_chunks,
                       uint32_t *__restrict__ summary) {
    __shared__ int sm_last[FP_THREADS / 32];
    const uint32_t chunk = blockIdx.x;
    if (chunk >= n_chunks) return;
    const FpFile f = files[fp_find_file(files, n_files, chunk)];
    const uint32_t lc = chunk - f.chunk_begin;
    const uint64_t base = f.text_off + (uint64_t)lc * FP_CHUNK;
    const uint32_t len = (uint32_t)min((uint64_t)FP_CHUNK, f.text_len - (uint64_t)lc * FP_CHUNK);
    int last = -1;
    const uint32_t o = min(threadIdx.x * FP_BYTES_PER_THREAD, len), e = min(o + FP_BYTES_PER_THREAD, len);
    for (uint32_t i = o; i < e; i++)
        if (text[base + i] == '\n') last = (int)i;
    for (int d = 16; d > 0; d >>= 1) last = max(last, __shfl_xor_sync(0xffffffffu, last, d));
    if ((threadIdx.x & 31) == 0) sm_last[threadIdx.x >> 5] = last;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 0; w < FP_THREADS / 32; w++) last = max(last, sm_last[w]);
        uint32_t code;
        if (last < 0) code = (len && text[base] == '>') ? FP_PASS_HDR : FP_PASS_SEQ;  // decides only after a LINE_START
        else if ((uint32_t)last + 1 == len) code = FP_SET_LINE_START;
        else code = (text[base + last + 1] == '>') ? FP_SET_HDR : FP_SET_SEQ;
        summary[chunk] = code;
    }
}

// ---- line state at the start of every chunk: one CTA per file walks its chunk summaries (parallel max-scan of the
// last state-setting chunk; the first pass-through chunk after a LINE_START decides by its first byte)
__global__ void __launch_bounds__(1024)
fp_line_state_kernel(const FpFile *__restrict__ files, const uint32_t *__restrict__ summary, uint32_t *__restrict__ state_in) {
    __shared__ int part[1024];
    const FpFile f = files[blockIdx.x];
    const uint32_t n = f.n_chunks;
    const uint32_t *sum = summary + f.chunk_begin;
    uint32_t *out = state_in + f.chunk_begin;
    const uint32_t per = (n + 1023) / 1024;
    const uint32_t b = min(threadIdx.x * per, n), e = min(b + per, n);
    int last = -1;  // last SET chunk in my range
    for (uint32_t i = b; i < e; i++)
        if (sum[i] >= FP_SET_SEQ) last = (int)i;
    part[threadIdx.x] = last;
    __syncthreads();
    if (threadIdx.x == 0) {
        int run = -1;
        for (int i = 0; i < 1024; i++) { const int t = part[i]; part[i] = run; run = max(run, t); }
    }
    __syncthreads();
    int prev = part[threadIdx.x];  // last SET chunk in front of my range
    for (uint32_t i = b; i < e; i++) {
        // state at the start of chunk i
        uint32_t s;
        if (prev < 0) {  // nothing set since the file start (= LINE_START): chunk 0 decided it if we are past it
            s = (i == 0) ? FP_ST_LINE_START : (sum[0] == FP_PASS_HDR ? FP_ST_HDR : FP_ST_SEQ);
        } else {
            const uint32_t code = sum[prev];
            if (code == FP_SET_SEQ) s = FP_ST_SEQ;
            else if (code == FP_SET_HDR) s = FP_ST_HDR;
            else s = ((uint32_t)prev + 1 == i) ? FP_ST_LINE_START : (sum[prev + 1] == FP_PASS_HDR ? FP_ST_HDR : FP_ST_SEQ);
        }
        out[i] = s;
        if (sum[i] >= FP_SET_SEQ) prev = (int)i;
    }
}

// the per-byte state machine shared by passes A2, B and D
#define FP_FOR_EACH_BYTE(th, sm_text, ON_BASE, ON_BREAK)                            \
    {                                                                               \
        bool hdr_ = (th).hdr, ls_ = (th).at_ls;                                     \
        for (uint32_t i_ = (th).o; i_ < (th).e; i_++) {                             \
            const uint8_t c_ = (sm_text)[i_];                                       \
            if (ls_ && c_ == '>') { hdr_ = true; ls_ = false; ON_BREAK }            \
            else if (hdr_) { if (c_ == '\n') { hdr_ = false; ls_ = true; } }        \
            else {                                                                  \
                const uint32_t v_ = fp_code(c_);                                    \
                if (v_ < 4) { ls_ = false; ON_BASE(v_) }                            \
                else if (v_ == 5) ls_ = true;                                       \
                else if (v_ == 4) ls_ = false;                                      \
                else { ls_ = false; ON_BREAK }                                      \
            }                                                                       \
        }                                                                           \
    }

__device__ __forceinline__ void fp_block_excl_scan2(uint32_t a, uint32_t b, uint32_t &ea, uint32_t &eb, uint32_t &ta, uint32_t &tb,
                                                    uint32_t *sm_a, uint32_t *sm_b) {
    uint32_t ia = a, ib = b;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t va = __shfl_up_sync(0xffffffffu, ia, d), vb = __shfl_up_sync(0xffffffffu, ib, d);
        if (lane >= d) { ia += va; ib += vb; }
    }
    __syncthreads();
    if (lane == 31) { sm_a[wid] = ia; sm_b[wid] = ib; }
    __syncthreads();
    uint32_t wa = 0, wb = 0;
    ta = tb = 0;
    for (int w = 0; w < FP_THREADS / 32; w++) {
        if (w < wid) { wa += sm_a[w]; wb += sm_b[w]; }
        ta += sm_a[w];
        tb += sm_b[w];
    }
    ea = wa + ia - a;
    eb = wb + ib - b;
}

// ---- A2: bases and break events per chunk
__global__ void __launch_bounds__(FP_THREADS)
fp_count_kernel(const uint8_t *__restrict__ text, const FpFile *__restrict__ files, uint32_t n_files, uint32_t n_chunks,
                const uint32_t *__restrict__ state_in, uint32_t *__restrict__ chunk_bases, uint32_t *__restrict__ chunk_breaks) {
    __shared__ uint8_t sm_text[FP_CHUNK];
    __shared__ int sm_scan[FP_THREADS / 32];
    __shared__ uint32_t sm_a[FP_THREADS / 32], sm_b[FP_THREADS / 32];
    const uint32_t chunk = blockIdx.x;
    const FpFile f = files[fp_find_file(files, n_files, chunk)];
    uint32_t len;
    const FpThread th = fp_stage(text, f, chunk - f.chunk_begin, state_in[chunk], sm_text, sm_scan, len);
    uint32_t nb = 0, nk = 0;
#define FP_ON_BASE(v) nb++;
    FP_FOR_EACH_BYTE(th, sm_text, FP_ON_BASE, nk++;)
#undef FP_ON_BASE
    uint32_t ea, eb, ta, tb;
    fp_block_excl_scan2(nb, nk, ea, eb, ta, tb, sm_a, sm_b);
    if (threadIdx.x == 0) { chunk_bases[chunk] = ta; chunk_breaks[chunk] = tb; }
}

// ---- per file: exclusive scan of its chunk counts in place, totals out (one CTA per file)
__global__ void __launch_bounds__(1024)
fp_chunk_scan_kernel(const FpFile *__restrict__ files, uint32_t *__restrict__ chunk_bases, uint32_t *__restrict__ chunk_breaks,
                     unsigned long long *__restrict__ totals /* [n_files][2] */) {
    __shared__ unsigned long long pa[1024], pb[1024];
    const FpFile f = files[blockIdx.x];
    uint32_t *ca = chunk_bases + f.chunk_begin, *cb = chunk_breaks + f.chunk_begin;
    const uint32_t n = f.n_chunks, per = (n + 1023) / 1024;
    const uint32_t b = min(threadIdx.x * per, n), e = min(b + per, n);
    unsigned long long sa = 0, sb = 0;
    for (uint32_t i = b; i < e; i++) { sa += ca[i]; sb += cb[i]; }
    pa[threadIdx.x] = sa;
    pb[threadIdx.x] = sb;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long ra = 0, rb = 0;
        for (int i = 0; i < 1024; i++) {
            const unsigned long long ta = pa[i], tb = pb[i];
            pa[i] = ra; pb[i] = rb;
            ra += ta; rb += tb;
        }
        totals[2 * blockIdx.x] = ra;
        totals[2 * blockIdx.x + 1] = rb;
    }
    __syncthreads();
    unsigned long long ra = pa[threadIdx.x], rb = pb[threadIdx.x];
    for (uint32_t i = b; i < e; i++) {
        const uint32_t ta = ca[i], tb = cb[i];
        ca[i] = (uint32_t)ra; cb[i] = (uint32_t)rb;
        ra += ta; rb += tb;
    }
}

// ---- B: break table
__global__ void __launch_bounds__(FP_THREADS)
fp_break_table_kernel(const uint8_t *__restrict__ text, const FpFile *__restrict__ files, uint32_t n_files, uint32_t n_chunks,
                      const uint32_t *__restrict__ state_in, const uint32_t *__restrict__ chunk_bases,
                      const uint32_t *__restrict__ chunk_breaks, uint32_t *__restrict__ bt) {
    __shared__ uint8_t sm_text[FP_CHUNK];
    __shared__ int sm_scan[FP_THREADS / 32];
    __shared__ uint32_t sm_a[FP_THREADS / 32], sm_b[FP_THREADS / 32];
    const uint32_t chunk = blockIdx.x;
    const FpFile f = files[fp_find_file(files, n_files, chunk)];
    uint32_t len;
    const FpThread th = fp_stage(text, f, chunk - f.chunk_begin, state_in[chunk], sm_text, sm_scan, len);
    uint32_t nb = 0, nk = 0;
#define FP_ON_BASE(v) nb++;
    FP_FOR_EACH_BYTE(th, sm_text, FP_ON_BASE, nk++;)
#undef FP_ON_BASE
    uint32_t ea, eb, ta, tb;
    fp_block_excl_scan2(nb, nk, ea, eb, ta, tb, sm_a, sm_b);
    if (nk) {
        uint32_t b_rank = chunk_bases[chunk] + ea, k_rank = chunk_breaks[chunk] + eb;
        uint32_t *out = bt + f.bt_off;
#define FP_ON_BASE(v) b_rank++;
        FP_FOR_EACH_BYTE(th, sm_text, FP_ON_BASE, out[k_rank++] = b_rank;)
#undef FP_ON_BASE
    }
}

// ---- C: per file, scan over its runs.  Run r = the bases between break events r-1 and r (r = 0 .. n_breaks).
// kp[r] = bases of kept runs in front of run r.  Writes run_ends and the file's totals (kept bases, kept runs).
__global__ void __launch_bounds__(1024)
fp_run_scan_kernel(const FpFile *__restrict__ files, const unsigned long long *__restrict__ totals, const uint32_t *__restrict__ bt,
                   uint32_t min_run, uint32_t *__restrict__ kp, uint32_t *__restrict__ run_ends,
                   unsigned long long *__restrict__ out_totals /* [n_files][2]: kept bases, kept runs */) {
    __shared__ unsigned long long pa[1024], pb[1024];
    const FpFile f = files[blockIdx.x];
    const uint32_t n_bases = (uint32_t)totals[2 * blockIdx.x];
    const uint32_t n_runs = (uint32_t)totals[2 * blockIdx.x + 1] + 1;
    const uint32_t *B = bt + f.bt_off;          // n_runs - 1 entries
    uint32_t *KP = kp + f.bt_off + blockIdx.x;  // n_runs entries per file (one more than the break table)
    const uint32_t need = max(min_run, 1u);
    auto len_of = [&](uint32_t r) -> uint32_t {
        const uint32_t hi = (r + 1 < n_runs) ? B[r] : n_bases, lo = r ? B[r - 1] : 0u;
        return hi - lo;
    };
    const uint32_t per = (n_runs + 1023) / 1024;
    const uint32_t b = min(threadIdx.x * per, n_runs), e = min(b + per, n_runs);
    unsigned long long sa = 0, sb = 0;
    for (uint32_t r = b; r < e; r++) {
        const uint32_t l = len_of(r);
        if (l >= need) { sa += l; sb++; }
    }
    pa[threadIdx.x] = sa;
    pb[threadIdx.x] = sb;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long ra = 0, rb = 0;
        for (int i = 0; i < 1024; i++) {
            const unsigned long long ta = pa[i], tb = pb[i];
            pa[i] = ra; pb[i] = rb;
            ra += ta; rb += tb;
        }
        out_totals[2 * blockIdx.x] = ra;
        out_totals[2 * blockIdx.x + 1] = rb;
    }
    __syncthreads();
    unsigned long long ra = pa[threadIdx.x], rb = pb[threadIdx.x];
    for (uint32_t r = b; r < e; r++) {
        const uint32_t l = len_of(r);
        KP[r] = (uint32_t)ra;
        if (l >= need) {
            ra += l;
            run_ends[f.run_off + rb] = (uint32_t)ra;
            rb++;
        }
    }
}

// ---- D: emit the kept bases
__global__ void __launch_bounds__(FP_THREADS)
fp_emit_kernel(const uint8_t *__restrict__ text, const FpFile *__restrict__ files, uint32_t n_files, uint32_t n_chunks,
               const uint32_t *__restrict__ state_in, const uint32_t *__restrict__ chunk_bases,
               const uint32_t *__restrict__ chunk_breaks, const unsigned long long *__restrict__ totals,
               const uint32_t *__restrict__ bt, const uint32_t *__restrict__ kp, uint32_t min_run,
               unsigned long long *__restrict__ packed) {
    __shared__ uint8_t sm_text[FP_CHUNK];
    __shared__ int sm_scan[FP_THREADS / 32];
    __shared__ uint32_t sm_a[FP_THREADS / 32], sm_b[FP_THREADS / 32];
    const uint32_t chunk = blockIdx.x;
    const uint32_t fi = fp_find_file(files, n_files, chunk);
    const FpFile f = files[fi];
    uint32_t len;
    const FpThread th = fp_stage(text, f, chunk - f.chunk_begin, state_in[chunk], sm_text, sm_scan, len);
    uint32_t nb = 0, nk = 0;
#define FP_ON_BASE(v) nb++;
    FP_FOR_EACH_BYTE(th, sm_text, FP_ON_BASE, nk++;)
#undef FP_ON_BASE
    uint32_t ea, eb, ta, tb;
    fp_block_excl_scan2(nb, nk, ea, eb, ta, tb, sm_a, sm_b);
    if (nb == 0) return;
    const uint32_t n_bases = (uint32_t)totals[2 * fi], n_runs = (uint32_t)totals[2 * fi + 1] + 1;
    const uint32_t *B = bt + f.bt_off;
    const uint32_t *KP = kp + f.bt_off + fi;
    const uint32_t need = max(min_run, 1u);
    uint32_t b_rank = chunk_bases[chunk] + ea, r = chunk_breaks[chunk] + eb;
    unsigned long long *out = packed + f.packed_word_off;
    // state of the current run: kept?, and (output position - base rank)
    bool kept = false;
    uint32_t delta = 0;
    auto load_run = [&]() {
        const uint32_t hi = (r + 1 < n_runs) ? B[r] : n_bases, lo = r ? B[r - 1] : 0u;
        kept = (hi - lo) >= need;
        delta = KP[r] - lo;
    };
    load_run();
    unsigned long long acc = 0;
    uint32_t acc_word = 0xFFFFFFFFu;
#define FP_ON_BASE(v)                                                                    \
    {                                                                                    \
        if (kept) {                                                                      \
            const uint32_t pos = b_rank + delta;                                         \
            const uint32_t w = pos >> 5;                                                 \
            if (w != acc_word) {                                                         \
                if (acc_word != 0xFFFFFFFFu) atomicOr(out + acc_word, acc);              \
                acc = 0;                                                                 \
                acc_word = w;                                                            \
            }                                                                            \
            acc |= (unsigned long long)(v) << (62 - 2 * (pos & 31));                     \
        }                                                                                \
        b_rank++;                                                                        \
    }
    FP_FOR_EACH_BYTE(th, sm_text, FP_ON_BASE, { r++; load_run(); })
#undef FP_ON_BASE
    if (acc_word != 0xFFFFFFFFu) atomicOr(out + acc_word, acc);
}
