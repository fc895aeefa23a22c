#!/usr/bin/env python
"""K1 across small filter sizes (real MetaSBT databases use 2^22..2^26 bits) and variants: us/genome.
    python tools/bench_small_filters.py [genomes] [genome_bp] [log2 sizes, comma separated] [variants, comma separated]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from metasbt_b200 import ops, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
bp = int(sys.argv[2]) if len(sys.argv) > 2 else 5_000_000
ctx = ops.Context(0)
dev = ctx.device
words = [synth.packed_genome(0x5EED0000 + g, bp, dev) for g in range(n)]
batch = ops.GenomeBatch.from_device_words(words, [bp] * n, dev)
sizes = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else list(range(20, 29))
variants = [int(x) for x in sys.argv[4].split(",")] if len(sys.argv) > 4 else [0, 1, 2, 3]
res = {}
for lb in sizes:
    out = ctx.alloc_filters(n, 1 << lb)
    row = {}
    for variant in variants:
        try:
            ctx.build_filters(batch, 31, 1 << lb, out=out, variant=variant)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ctx.build_filters(batch, 31, 1 << lb, out=out, variant=variant)
            e1.record()
            torch.cuda.synchronize()
            row["v%d" % variant] = round(e0.elapsed_time(e1) * 1e3 / n, 1)
        except Exception as exc:
            row["v%d" % variant] = "n/a"
    res["2^%d" % lb] = row
    del out
print(json.dumps({"lib": os.environ.get("MSBT_LIB", "default"), "direct_group_mb": os.environ.get("MSBT_DIRECT_GROUP_MB"), "genomes": n, "genome_bp": bp, "us_per_genome": res}))
