#!/usr/bin/env python
"""k-mer size search (kopt.py; the role of `kitsune kopt` in core.py:439-468) on synthetic related genomes: wall time of
the three steps from FASTA files on disk and the chosen k.  `species` x `per_species` genomes of `bp` bases, each a 1-3 %
mutant of its species ancestor.
    python tools/bench_kopt.py [species] [per_species] [bp] [k_max]"""
import json, os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from metasbt_b200 import kopt
from metasbt_b200.backend import CudaBackend
from tests import util

species = int(sys.argv[1]) if len(sys.argv) > 1 else 4
per = int(sys.argv[2]) if len(sys.argv) > 2 else 4
bp = int(sys.argv[3]) if len(sys.argv) > 3 else 5_000_000
k_max = int(sys.argv[4]) if len(sys.argv) > 4 else 32
d = tempfile.mkdtemp()
paths = []
for s in range(species):
    anc = util.synthetic_codes(0xA11CE000 + s, bp)
    for i in range(per):
        p = os.path.join(d, "s%d_g%d.fna" % (s, i))
        with open(p, "wb") as f:
            f.write(util.codes_to_fasta(util.mutate(anc, 0.01 * (1 + i % 3), 1000 + s * per + i), "s%d_g%d" % (s, i), 80))
        paths.append(p)
be = CudaBackend(k_max, kopt.SET_BITS)
kopt.optimal_kmer_size(paths[:2], min(k_max, 8), be)  # warm-up
torch.cuda.synchronize()
launches = be.ctx.launch_count
t0 = time.perf_counter()
codes = [kopt.fasta_codes(be.read_fasta(p)) for p in paths]
t1 = time.perf_counter()
for c in codes:
    kopt.relative_entropies(c, k_max, be.device)
torch.cuda.synchronize()
t2 = time.perf_counter()
details = {}
k = kopt.optimal_kmer_size(paths, k_max, be, details=details)
torch.cuda.synchronize()
t3 = time.perf_counter()
print(json.dumps({"genomes": len(paths), "genome_bp": bp, "k_max": k_max, "k": k, "k_lo": details["k_lo"], "k_hi": details.get("k_hi"),
                  "cre_limits": details["cre_limits"], "sample": details.get("sample"), "set_filter_bits": kopt.SET_BITS,
                  "seconds_parse_host": round(t1 - t0, 3), "seconds_word_counting": round(t2 - t1, 3),
                  "seconds_whole_search": round(t3 - t2, 3), "kernel_launches": be.ctx.launch_count - launches,
                  "acf": {str(kk): round(v) for kk, v in details.get("acf", {}).items()},
                  "ofc": {str(kk): round(v) for kk, v in details.get("ofc", {}).items()},
                  "note": "whole search = host parse + word counting (step 1) + per k: device parse/pack, K1, K3 all pairs, occurrence pass"}))
