#!/usr/bin/env python
"""BASELINE.json configs[4]: index + profile sweep k in {21, 31, 51} x filter sizes 2^28..2^32 on one GPU -- prints one JSON
object.  Index leg: K1 over `--genomes` 5 Mbp genomes against the HBM roofline (algorithmic bytes ceil(G/4) + m/8 per genome).
Profile leg: `--mags` MAGs (2 Mbp, 3 % mutants of the indexed genomes) hashed (K1), their distinct positions extracted and probed
against the indexed genomes as one flat tree (row-major probe: one 32-byte sector per position and leaf).
    python tools/sweep.py [--genomes 64] [--mags 128] [--variant 0]"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from metasbt_b200 import ops, synth

ap = argparse.ArgumentParser()
ap.add_argument("--genomes", type=int, default=64)
ap.add_argument("--genome-bp", type=int, default=5_000_000)
ap.add_argument("--variant", type=int, default=0)
ap.add_argument("--mags", type=int, default=128)
ap.add_argument("--mag-bp", type=int, default=2_000_000)
a = ap.parse_args()
peak = 6537.6
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
ctx = ops.Context(0)
dev = ctx.device
G = a.genome_bp
res = {"genome_bp": G, "variant": a.variant, "hbm_peak_gbs": peak, "rows": []}
for logm in (28, 29, 30, 31, 32):
    m = 1 << logm
    n = min(a.genomes, max(8, int(100e9 // (m // 8))))
    words = [synth.packed_genome(0x5EED0000 + g, G, dev) for g in range(n)]
    batch = ops.GenomeBatch.from_device_words(words, [G] * n, dev)
    del words
    out = ctx.alloc_filters(n, m)
    for k in (21, 31, 51):
        ctx.build_filters(batch, k, m, out=out, variant=a.variant)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            ctx.build_filters(batch, k, m, out=out, variant=a.variant)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 3 / n * 1e3
        algo = (G + 3) // 4 + m // 8
        row = {"k": k, "log2_bits": logm, "genomes": n, "us_per_genome": us, "Mbp_per_s": G / us,
               "GBps": algo / us / 1e3, "frac_of_hbm_peak": algo / us / 1e3 / peak}
        # profile leg: the filters just built are the leaves of one flat tree
        nq = min(a.mags, max(8, int(40e9 // (m // 8))))
        qwords = [synth.pack_codes_device(synth.mutated_codes(0x5EED0000 + (q * 7) % n, 0xBEEF0000 + q, a.mag_bp, 30, dev)) for q in range(nq)]
        # (the leaves are plain synthetic genomes seeded 0x5EED0000 + g: a MAG is a 3 % mutant of a prefix of leaf (7q mod n))
        qbatch = ops.GenomeBatch.from_device_words(qwords, [a.mag_bp] * nq, dev)
        del qwords
        qf = ctx.alloc_filters(nq, m)
        leaves = [out[i] for i in range(n)]

        def timed(fn):
            fn()
            torch.cuda.synchronize()
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
            r = fn()
            t1.record()
            torch.cuda.synchronize()
            return r, t0.elapsed_time(t1)
        _, ms_hash = timed(lambda: ctx.build_filters(qbatch, k, m, out=qf, variant=a.variant))
        (pos, offs), ms_extract = timed(lambda: ctx.extract_positions_packed([qf[i] for i in range(nq)], m, cap=nq * a.mag_bp))
        ranges = [(offs[i], offs[i + 1]) for i in range(nq)]
        hits, ms_probe = timed(lambda: ctx.query_ranges(leaves, pos, ranges))
        best = hits.argmax(dim=1).cpu().tolist()
        n_pos = int(offs[-1])
        row["profile"] = {"mags": nq, "leaves": n, "ms_hash": ms_hash, "ms_extract": ms_extract, "ms_probe": ms_probe,
                          "mags_per_s": nq / ((ms_hash + ms_extract + ms_probe) / 1e3), "positions": n_pos,
                          "probe_GBps_sectors": n_pos * n * 32 / ms_probe / 1e6, "probe_frac_of_hbm_peak": n_pos * n * 32 / ms_probe / 1e6 / peak,
                          "best_hit_is_source": sum(1 for q in range(nq) if best[q] == (q * 7) % n)}
        res["rows"].append(row)
        del qf, qbatch, pos, hits
    del out, batch
    torch.cuda.empty_cache()
print(json.dumps(res))
