#!/usr/bin/env python
"""Kernel-tuning aid: run the fused K1 kernel from a -DMSBT_FU_TRACE build and print CTA 0's per-phase clock totals.
    MSBT_LIB=$PWD/metasbt_b200/libmsbt_trace.so python tools/fu_trace.py [genomes]"""
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
ctx.build_filters(batch, 31, 1 << 30, out=out, variant=3); torch.cuda.synchronize()
L.msbt_debug_trace(buf, 1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); ctx.build_filters(batch, 31, 1 << 30, out=out, variant=3); e1.record(); torch.cuda.synchronize()
L.msbt_debug_trace(buf, 0)
ms = e0.elapsed_time(e1)
print("total %.1f us/genome" % (ms * 1e3 / n))
print("flush (warp 0): reservations done after %.2f us, copy done after %.2f us per chunk" % (buf[5] / 1.965e3 / max(buf[7], 1), buf[6] / 1.965e3 / max(buf[7], 1)))
names = ["P barA (wait claim)", "P hash+barB", "(unused)", "P flush+fence", "P ring wait/claim", "", "", "P chunks",
         "B wait_group.read (tid0)", "B barrier after wait", "B zero+select+bar", "B atomics", "B fence+bar", "B claim wait+bar", "B issue loads", "B tiles",
         "B store issue (tid0)", "B loop top (tid0)", "B loop top (tid32)"]
clk = 1.965e3  # cycles per us at max clock (approx.)
for i, nm in enumerate(names):
    if nm:
        v = buf[i]
        if nm.endswith("chunks") or nm.endswith("tiles"):
            print("%-28s %d" % (nm, v))
        else:
            cnt = buf[7] if i < 8 else buf[15]
            print("%-28s total %8.1f us   per-unit %7.2f us" % (nm, v / clk, v / clk / max(cnt, 1)))
