#!/usr/bin/env python
"""BASELINE.json configs[4]: K1 throughput sweep k in {21, 31, 51} x filter sizes 2^28..2^32 on one GPU, each against the
HBM roofline (algorithmic bytes ceil(G/4) + m/8 per genome) -- prints one JSON object.
    python tools/sweep.py [--genomes 64] [--variant 0]"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from metasbt_b200 import ops, synth

ap = argparse.ArgumentParser()
ap.add_argument("--genomes", type=int, default=64)
ap.add_argument("--genome-bp", type=int, default=5_000_000)
ap.add_argument("--variant", type=int, default=0)
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
        res["rows"].append({"k": k, "log2_bits": logm, "genomes": n, "us_per_genome": us, "Mbp_per_s": G / us,
                            "GBps": algo / us / 1e3, "frac_of_hbm_peak": algo / us / 1e3 / peak})
    del out, batch
    torch.cuda.empty_cache()
print(json.dumps(res))
