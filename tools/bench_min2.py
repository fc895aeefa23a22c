#!/usr/bin/env python
"""K1 with --min-kmer-occurrences >= 2 (MetaSBT's CLI default is 2): time the exact-multiplicity build for 5 Mbp genomes.
    python tools/bench_min2.py [genomes] [log2 bits]        MSBT_MIN_TABLE=1 selects the plain count tables"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from metasbt_b200 import ops, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
lb = int(sys.argv[2]) if len(sys.argv) > 2 else 30
ctx = ops.Context(0)
dev = ctx.device
bp = 5_000_000
words = [synth.packed_genome(0x5EED0000 + g, bp, dev) for g in range(n)]
batch = ops.GenomeBatch.from_device_words(words, [bp] * n, dev)
out = ctx.alloc_filters(n, 1 << lb)
res = {}
for mn in (1, 2, 3):
    ctx.build_filters(batch, 31, 1 << lb, min_abund=mn, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ctx.build_filters(batch, 31, 1 << lb, min_abund=mn, out=out)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / n
    res["min%d" % mn] = {"us_per_genome": round(us, 1), "Gbp_per_s": round(bp / us / 1e3, 2), "popcount_filter0": int(ctx.bf_popcount([out[0]], 1 << lb)[0])}
print(json.dumps({"genomes": n, "log2_bits": lb, "path": "tables" if os.environ.get("MSBT_MIN_TABLE") else "prefiltered", **res}))
