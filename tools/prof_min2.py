#!/usr/bin/env python
"""One `--min 2` build of a few 5 Mbp genomes, for an ncu launch list.   python tools/prof_min2.py [genomes]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from metasbt_b200 import ops, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
ctx = ops.Context(0)
words = [synth.packed_genome(0x5EED0000 + g, 5_000_000, ctx.device) for g in range(n)]
batch = ops.GenomeBatch.from_device_words(words, [5_000_000] * n, ctx.device)
out = ctx.alloc_filters(n, 1 << 30)
ctx.build_filters(batch, 31, 1 << 30, min_abund=2, out=out)
torch.cuda.synchronize()
