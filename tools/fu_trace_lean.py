#!/usr/bin/env python
"""Kernel-tuning aid: per-phase clock totals of CTA 0's lean B loop (fused K1 kernel, -DMSBT_FU_TRACE build).
    MSBT_LIB=$PWD/metasbt_b200/libmsbt_trace.so python tools/fu_trace_lean.py [genomes] [log2 bits]"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from metasbt_b200 import _lib, ops, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
lb = int(sys.argv[2]) if len(sys.argv) > 2 else 30
ctx = ops.Context(0)
dev = ctx.device
words = [synth.packed_genome(0x5EED0000 + g, 5_000_000, dev) for g in range(n)]
batch = ops.GenomeBatch.from_device_words(words, [5_000_000] * n, dev)
out = ctx.alloc_filters(n, 1 << lb)
L = _lib.lib()
buf = (ctypes.c_ulonglong * 24)()
ctx.build_filters(batch, 31, 1 << lb, out=out, variant=3); torch.cuda.synchronize()
L.msbt_debug_trace(buf, 1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); ctx.build_filters(batch, 31, 1 << lb, out=out, variant=3); e1.record(); torch.cuda.synchronize()
L.msbt_debug_trace(buf, 0)
print("total %.1f us/genome" % (e0.elapsed_time(e1) * 1e3 / n))
clk = 1.965e3
tiles = max(buf[15], 1)
names = {10: "A zero + next-claim (tid 0)", 9: "barrier after A (tid 96)", 11: "B REDs (tid 0)", 12: "fence + barrier after B (tid 96)",
         14: "C prefetch issue (tid 0)", 8: "C bulk store + wait read-out (tid 0)", 13: "barrier after C (tid 96)",
         16: "whole tile (tid 0)", 17: "whole tile (tid 96)"}
print("tiles built by CTA 0: %d  (%.1f per genome)" % (tiles, tiles / n))
for i in (10, 9, 11, 12, 14, 8, 13, 16, 17):
    print("%-40s per tile %6.2f us" % (names[i], buf[i] / clk / tiles))
print("P: chunks %d; hash+barB %.2f us, flush+fence %.2f us, ring wait/claim %.2f us, barA %.2f us per chunk" % (
    buf[7], buf[1] / clk / max(buf[7], 1), buf[3] / clk / max(buf[7], 1), buf[4] / clk / max(buf[7], 1), buf[0] / clk / max(buf[7], 1)))
