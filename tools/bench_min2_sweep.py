#!/usr/bin/env python
"""Tuning sweep of the `--min 2` count pass (candidate bitmap in shared memory): bitmap size, work-unit size, probe hash.
The switches are read by the library on every call, so one process walks all settings.
    python tools/bench_min2_sweep.py [genomes] [log2 bits]"""
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


def timed(env):
    for key in ("MSBT_PRE_COUNT_CB", "MSBT_PRE_CB_LOG2", "MSBT_PRE_CB_UNIT", "MSBT_PRE_CB_MULT", "MSBT_PRE_STREAMS", "MSBT_PRE_ZERO_INLINE", "MSBT_PRE_BITS_PER_BASE"):
        os.environ.pop(key, None)
    os.environ.update(env)
    ctx.build_filters(batch, 31, 1 << lb, min_abund=2, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        ctx.build_filters(batch, 31, 1 << lb, min_abund=2, out=out)
    e1.record()
    torch.cuda.synchronize()
    return round(e0.elapsed_time(e1) * 1e3 / n / 3, 1), int(ctx.bf_popcount([out[1]], 1 << lb)[0])


rows = []
settings = [{}]
for bpb in ("32", "16", "8", "4"):
    for log2 in ("19", "20"):
        settings.append({"MSBT_PRE_BITS_PER_BASE": bpb, "MSBT_PRE_CB_LOG2": log2})
settings.append({"MSBT_PRE_BITS_PER_BASE": "16", "MSBT_PRE_CB_LOG2": "20", "MSBT_PRE_STREAMS": "4"})
for env in settings:
    us, pop = timed(env)
    rows.append({"env": env, "us_per_genome": us, "popcount_filter1": pop})
    print(rows[-1], file=sys.stderr)
print(json.dumps({"genomes": n, "log2_bits": lb, "genome_bp": bp, "rows": rows}))
