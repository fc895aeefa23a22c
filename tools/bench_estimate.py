#!/usr/bin/env python
"""Filter-size estimation (estimate.py; the role of ntCard in core.py:470-529) on synthetic 5 Mbp genomes:
wall time of backend.count_union_bits from FASTA files on disk, and the estimate against the known k-mer count.
    python tools/bench_estimate.py [files]"""
import json, os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from metasbt_b200 import estimate
from metasbt_b200.backend import CudaBackend
from tests import util

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
bp, k = 5_000_000, 31
d = tempfile.mkdtemp()
paths = []
for g in range(n):  # unrelated random genomes: practically all k-mers distinct
    p = os.path.join(d, "g%d.fna" % g)
    with open(p, "wb") as f:
        f.write(util.codes_to_fasta(util.synthetic_codes(0x5EED0000 + g, bp), "g%d" % g, 80))
    paths.append(p)
be = CudaBackend(k, estimate.EST_BITS)
be.count_union_bits(paths[:2], k, estimate.EST_BITS, estimate.EST_BITS)  # warm-up
torch.cuda.synchronize()
t0 = time.perf_counter()
f0, s = estimate.estimate_distinct_kmers(paths, k, be.count_union_bits)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
true = n * (bp - k + 1)
print(json.dumps({"files": n, "genome_bp": bp, "kmer_size": k, "counting_filter_bits": estimate.EST_BITS, "sampling_factor": s,
                  "seconds": round(dt, 4), "genomes_per_s": round(n / dt, 1), "Mbp_per_s": round(n * bp / dt / 1e6, 1),
                  "F0_estimate": f0, "kmers_in_input": true, "rel_diff": round((f0 - true) / true, 6),
                  "note": "wall clock from FASTA files on disk: read + H2D + device parse/pack + K1 (shared filter) + OR + popcount"}))
