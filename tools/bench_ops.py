#!/usr/bin/env python
"""Secondary measurements for the other rows of SURVEY.md 8a (K2 union, K3/K3' popcounts, K4 query) on one GPU.
Used by bench.py (key "secondary" of its JSON line) and runnable alone:

    python tools/bench_ops.py [--leaves 1024] [--filter-bits 268435456] [--mags 64]

Every number is timed with CUDA events on the launching stream after a warm-up call; inputs are far larger
than L2.  `frac` = algorithmic bytes / time / measured HBM peak.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _time(fn, reps=3):
    import torch
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3  # seconds


def run(ctx, n_leaves=1024, filter_bits=1 << 28, n_mags=64, genome_bp=5_000_000, mag_bp=2_000_000, k=31, peak_gbs=6537.6):
    import torch
    from metasbt_b200 import ops, synth
    dev = ctx.device
    m, L = filter_bits, n_leaves
    fbytes = m // 8
    out = {"config": {"leaves": L, "filter_bits": m, "mags": n_mags, "genome_bp": genome_bp, "mag_bp": mag_bp, "kmer_size": k}}

    def entry(seconds, algo_bytes, **extra):
        d = {"ms": seconds * 1e3, "algorithmic_GB": algo_bytes / 1e9, "GBps": algo_bytes / seconds / 1e9,
             "frac_of_hbm_peak": algo_bytes / seconds / 1e9 / peak_gbs}
        d.update(extra)
        return d

    # leaves: related genomes (species ancestor + 1 % SNPs), so that similarities and hits are non-trivial
    words = []
    for g in range(L):
        codes = synth.genome_codes(0xA11CE000 + g // 8, genome_bp, dev)
        r = synth.splitmix64(((0x5EED0000 + g) << 32) ^ torch.arange(genome_bp, dtype=torch.int64, device=dev))
        snp = (synth._lsr(r, 8) % 1000) < 10
        codes = torch.where(snp, (codes + 1 + synth._lsr(r, 40) % 3) % 4, codes)
        words.append(synth.pack_codes_device(codes))
    batch = ops.GenomeBatch.from_device_words(words, [genome_bp] * L, dev)
    del words
    leaves = ctx.alloc_filters(L, m)
    t = _time(lambda: ctx.build_filters(batch, k, m, out=leaves), reps=2)
    out["K1_build_leaves"] = entry(t, L * ((genome_bp + 3) // 4 + fbytes), Mbp_per_s=L * genome_bp / t / 1e6)
    rows = [leaves[i] for i in range(L)]
    # K1 with --min-kmer-occurrences 2 (MetaSBT's CLI default): exact multiplicity path, a bounded sample of the batch
    n2 = min(32, L)
    sub = ops.GenomeBatch.from_device_words([batch.d_words[i * (batch.d_words.numel() // L):(i + 1) * (batch.d_words.numel() // L)] for i in range(n2)],
                                            [genome_bp] * n2, dev)
    tmp = ctx.alloc_filters(n2, m)
    t2 = _time(lambda: ctx.build_filters(sub, k, m, min_abund=2, out=tmp), reps=1)
    out["K1_build_min2_%d" % n2] = entry(t2, n2 * ((genome_bp + 3) // 4 + fbytes), Mbp_per_s=n2 * genome_bp / t2 / 1e6,
                                         note="exact canonical k-mer multiplicity per genome (hash table), not a bandwidth-roofline path")
    del tmp, sub

    # N1: FASTA text -> packed genomes on the device (the stage in front of K1), from host bytes, H2D included
    try:
        from tests import util
        n_fa = 32
        texts = [util.codes_to_fasta(util.synthetic_codes(0x5EED0000 + g, genome_bp), "g%d" % g, 80) for g in range(4)]
        texts = [texts[i % 4] for i in range(n_fa)]
        nbytes = sum(len(t) for t in texts)
        import time as _time_mod
        ctx.pack_fasta_device(texts, k)
        torch.cuda.synchronize()
        t0 = _time_mod.perf_counter()
        fb = ctx.pack_fasta_device(texts, k)
        torch.cuda.synchronize()
        dt = _time_mod.perf_counter() - t0
        out["N1_fasta_ingest_device_%d" % n_fa] = {"ms": dt * 1e3, "text_GB": nbytes / 1e9, "GBps_text": nbytes / dt / 1e9,
                                                   "Mbp_per_s": fb.total_bases / dt / 1e6, "genomes_per_s": n_fa / dt,
                                                   "note": "wall clock from host bytes: copy into pinned memory + H2D + 7 kernels + 2 syncs"}
        t0 = _time_mod.perf_counter()
        ops.pack_fasta(texts[0], k)
        dth = _time_mod.perf_counter() - t0
        out["N1_fasta_ingest_host_1thread"] = {"ms_per_genome": dth * 1e3, "Mbp_per_s": genome_bp / dth / 1e6}
        del fb
    except Exception as exc:
        out["N1_fasta_ingest_device"] = {"error": repr(exc)}

    # K2: cluster filter = OR of C children
    C = min(100, L)
    t = _time(lambda: ctx.bf_or(rows[:C], m))
    out["K2_or_%d" % C] = entry(t, (C + 1) * fbytes)
    # K3': density numerators of all leaves
    t = _time(lambda: ctx.bf_popcount(rows, m))
    out["K3p_popcount_%d" % L] = entry(t, L * fbytes)
    # K3: one focus against all others (intersection and union in one pass)
    t = _time(lambda: ctx.bf_dist(rows[0], rows[1:], m))
    out["K3_dist_1v%d" % (L - 1)] = entry(t, L * fbytes, pairs_per_s=(L - 1) / t)
    # K3 all pairs inside one species-sized cluster
    A = min(64, L)
    t = _time(lambda: ctx.bf_allpairs(rows[:A], m))
    out["K3_allpairs_%d" % A] = entry(t, A * fbytes, pairs_per_s=A * (A - 1) / 2 / t,
                                      note="bytes = each filter read once (lower bound); the kernel re-reads tiles")

    # ... and inside a 10-genome species (the shape Entry.get_boundaries sees most often)
    t = _time(lambda: ctx.bf_allpairs(rows[:10], m))
    out["K3_allpairs_10"] = entry(t, 10 * fbytes, pairs_per_s=45 / t, note="bytes = each filter read once")

    # ... and over ALL leaves at once: K3 (AND + popcount per pair) against the K4 route (transpose once, every filter's set
    # bits probe the bit-sliced matrix; ops.bf_allpairs_probe) -- same integers
    if L >= 256:
        box = {}
        t_bit = _time(lambda: box.__setitem__("bit", ctx.bf_allpairs(rows, m)), reps=1)
        t_probe = _time(lambda: box.__setitem__("probe", ctx.bf_allpairs_probe(rows, m)), reps=1)
        pairs = L * (L - 1) / 2
        out["K3_allpairs_%d_bit_route" % L] = {"ms": t_bit * 1e3, "pairs_per_s": pairs / t_bit}
        out["K3_allpairs_%d_probe_route" % L] = {"ms": t_probe * 1e3, "pairs_per_s": pairs / t_probe, "speedup": t_bit / t_probe,
                                                 "equal_to_bit_route": bool(torch.equal(box["bit"], box["probe"])),
                                                 "note": "work ~ set bits x row width instead of n^2 / 2 x filter words"}
        del box

    # K4: MAGs = 5 % mutated prefixes of leaves
    mw = []
    for q in range(n_mags):
        g = (q * 37) % L
        codes = synth.genome_codes(0xA11CE000 + g // 8, mag_bp, dev)
        r = synth.splitmix64(((0xBEEF0000 + q) << 32) ^ torch.arange(mag_bp, dtype=torch.int64, device=dev))
        snp = (synth._lsr(r, 8) % 1000) < 50
        codes = torch.where(snp, (codes + 1 + synth._lsr(r, 40) % 3) % 4, codes)
        mw.append(synth.pack_codes_device(codes))
    qbatch = ops.GenomeBatch.from_device_words(mw, [mag_bp] * n_mags, dev)
    del mw
    qf = ctx.alloc_filters(n_mags, m)
    t_sk = _time(lambda: ctx.build_filters(qbatch, k, m, out=qf), reps=2)
    qrows = [qf[i] for i in range(n_mags)]
    positions = ctx.extract_positions_batch(qrows, m, cap=n_mags * mag_bp)
    t_ex = _time(lambda: ctx.extract_positions_batch(qrows, m, cap=n_mags * mag_bp), reps=2)
    n_pos = sum(int(p.numel()) for p in positions)
    out["K4_query_hash_and_extract"] = {"ms_hash": t_sk * 1e3, "ms_extract": t_ex * 1e3, "positions": n_pos,
                                        "algorithmic_GB_extract": 2 * n_mags * fbytes / 1e9, "GBps_extract": 2 * n_mags * fbytes / t_ex / 1e9,
                                        "note": "distinct positions = set bits of the MAG's own min=1 filter (K1 + batched ordered extraction, 2 passes over the filters)"}
    t_bs = _time(lambda: ctx.bitslice(rows, 0, m), reps=1)
    sliced = ctx.bitslice(rows, 0, m)
    out["K4_bitslice_build"] = entry(t_bs, 2 * L * fbytes, note="transpose: read every leaf once, write the sliced matrix once")
    hits = ctx.query_batch_sliced(sliced, L, 0, m, positions)
    t_q = _time(lambda: ctx.query_batch_sliced(sliced, L, 0, m, positions), reps=2)
    algo = n_pos * ((L + 7) // 8) + n_mags * 4 * L
    out["K4_query_sliced"] = entry(t_q, algo, mags_per_s=n_mags / t_q,
                                   mags_per_s_incl_hash_extract=n_mags / (t_q + t_sk + t_ex))
    # cross-check against K3 (App. A6 corollary: hits == popc(Q & F)) on the first MAG
    inter, _ = ctx.bf_dist(qf[0], rows, m)
    out["K4_check_hits_equal_and_popcount"] = bool(torch.equal(hits[0].to(torch.int64), inter))
    # row-major probe on a species-sized leaf set (what `profile` uses inside one cluster)
    S = min(16, L)
    t_r = _time(lambda: ctx.query_batch(rows[:S], positions[:8]), reps=2)
    out["K4_query_rowmajor_8x%d" % S] = entry(t_r, sum(int(p.numel()) for p in positions[:8]) * S * 32,
                                              note="bytes = one 32-byte sector per probe")
    del sliced, leaves, rows
    torch.cuda.empty_cache()
    # BASELINE configs[3] shape on ONE bit-range shard: 10,000 leaves, rows [0, m/8) of a 2^28-bit filter space
    # (what each of 8 GPUs holds).  The leaf bits are random words (building 10,000 real leaves needs 8 GPUs);
    # the queries are the real MAG positions above, so the gather pattern and the traffic are the real ones.
    try:
        L4, shard_rows = 10_000, m // 8
        lw4 = ctx.sliced_row_words(L4)
        big = torch.empty((shard_rows, lw4), dtype=torch.int32, device=dev)
        step = 1 << 20
        gen = torch.Generator(device=dev)
        gen.manual_seed(1)
        for r0 in range(0, shard_rows, step):
            big[r0:r0 + step] = torch.randint(-2 ** 31, 2 ** 31 - 1, (min(step, shard_rows - r0), lw4), dtype=torch.int32, device=dev, generator=gen)
        h4 = ctx.query_batch_sliced(big, L4, 0, shard_rows, positions)
        t4 = _time(lambda: ctx.query_batch_sliced(big, L4, 0, shard_rows, positions), reps=2)
        n_in = sum(int((p.view(torch.int32) >= 0).logical_and(p < shard_rows).sum()) for p in positions)
        algo4 = n_in * ((L4 + 7) // 8) + n_mags * 4 * L4
        # spot check of one (query, leaf) count against a direct gather
        p0 = positions[0]
        p0 = p0[(p0 >= 0) & (p0 < shard_rows)].long()
        want = int(((big[p0, 7] >> 3) & 1).sum())
        out["K4_query_sliced_10000_leaves_one_of_8_shards"] = entry(
            t4, algo4, mags_per_s_per_shard=n_mags / t4, positions_in_shard=n_in, spot_check=bool(int(h4[0, 7 * 32 + 3]) == want),
            note="row = 1,250 B per position; with 8 GPUs each holding one shard the batch rate is this rate (partial hits are summed by all_reduce)")
        del big, h4
    except Exception as exc:
        out["K4_query_sliced_10000_leaves_one_of_8_shards"] = {"error": repr(exc)}
    del qf
    torch.cuda.empty_cache()
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--leaves", type=int, default=1024)
    ap.add_argument("--filter-bits", type=int, default=1 << 28)
    ap.add_argument("--mags", type=int, default=64)
    a = ap.parse_args()
    import torch
    from metasbt_b200 import ops
    peak = 6537.6
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    print(json.dumps(run(ops.Context(0), a.leaves, a.filter_bits, a.mags, peak_gbs=peak)))
