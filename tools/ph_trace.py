#!/usr/bin/env python
"""Kernel-tuning aid: per-phase clock totals of CTA 0 of the phased K1 kernel (-DMSBT_FU_TRACE build).
    MSBT_LIB=$PWD/metasbt_b200/libmsbt_trace.so python tools/ph_trace.py [genomes]"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from metasbt_b200 import _lib, ops, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
ctx = ops.Context(0)
dev = ctx.device
words = [synth.packed_genome(0x5EED0000 + g, 5_000_000, dev) for g in range(n)]
batch = ops.GenomeBatch.from_device_words(words, [5_000_000] * n, dev)
out = ctx.alloc_filters(n, 1 << 30)
L = _lib.lib()
buf = (ctypes.c_ulonglong * 24)()
ctx.build_filters(batch, 31, 1 << 30, out=out, variant=4); torch.cuda.synchronize()
L.msbt_debug_trace(buf, 1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); ctx.build_filters(batch, 31, 1 << 30, out=out, variant=4); e1.record(); torch.cuda.synchronize()
L.msbt_debug_trace(buf, 0)
ms = e0.elapsed_time(e1)
print("total %.1f us/genome, kernel %.0f us" % (ms * 1e3 / n, ms * 1e3))
clk = 1.965e3
rows = [(0, "decide + barrier", None), (1, "HASH: wait stores", 4), (2, "HASH: hash loop", 4), (3, "HASH: flush+publish", 4),
        (8, "BUILD: prologue", 13), (9, "tile: wait_group.read", 14), (10, "tile: S1 barrier", 14), (11, "tile: REDs+zero+S2", 14),
        (12, "tile: store issue", 14)]
print("spin polls: ring-blocked only %d, item-blocked only %d, both %d" % (buf[17], buf[18], buf[19]))
print("flush: atomics issued+returned (tid0) %.2f us, through copy %.2f us per chunk" % (buf[5] / clk / max(buf[4], 1), buf[6] / clk / max(buf[4], 1)))
print("HASH phases %d, BUILD phases %d, tiles %d" % (buf[4], buf[13], buf[14]))
for i, nm, c in rows:
    cnt = (buf[4] + buf[13]) if c is None else buf[c]
    print("%-26s total %8.1f us   per-unit %7.2f us" % (nm, buf[i] / clk, buf[i] / clk / max(cnt, 1)))
