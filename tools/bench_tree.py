#!/usr/bin/env python
"""BASELINE.json configs[2]: SBT tree build (filter union + pairwise popcount similarity) over synthetic genomes,
filters resident in HBM, on 1..N GPUs of one box (one process per GPU):

    python tools/bench_tree.py [--genomes 2000] [--filter-bits 268435456]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/bench_tree.py ...

What is timed is what `Database.update()` + `_dump_report()` do on the reference (core.py:1388-1459, 1503-1662),
stage by stage, without the file I/O:
  K1  leaf filters (makebf)                                     -- genomes sharded, no collective
  K2  bottom-up unions: species <- genomes, genus <- species ... kingdom <- phyla, db root <- kingdoms
  K3  all-pairs intersections inside every species (get_boundaries) and among the species centroids of every
      higher cluster, plus density popcounts of every cluster filter
The taxonomy is a balanced 7-level tree: 10 genomes per species, then `fanout` children per level.  Whole phyla are
assigned to ranks, so only the two top unions (kingdom, db root) cross GPUs: their partial filters are merged by
bit range (all_to_all + local OR + all_gather, shard.or_merge) because NCCL has no bitwise-OR reduction.
Prints one JSON object (rank 0): per-stage seconds (max over ranks), bytes and fraction of the HBM roofline.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

from bench import ClockSampler
from metasbt_b200 import ops, shard, synth


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--genomes", type=int, default=2000, help="total genomes (all ranks)")
    ap.add_argument("--genome-bp", type=int, default=5_000_000)
    ap.add_argument("--filter-bits", type=int, default=1 << 28)
    ap.add_argument("--kmer", type=int, default=31)
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=dev)
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(saved)
    ctx = ops.Context(local)
    k, m, G = a.kmer, a.filter_bits, a.genome_bp
    fbytes = m // 8
    peak = 6537.6
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass

    # taxonomy: species of 10 genomes; 8 phyla at the top (one or more per rank); balanced in between
    n_species = a.genomes // 10
    n_phyla = max(8, world)
    sp_per_phylum = max(1, n_species // n_phyla)
    n_species = sp_per_phylum * n_phyla
    my_phyla = [p for p in range(n_phyla) if p % world == rank]
    my_species = [p * sp_per_phylum + s for p in my_phyla for s in range(sp_per_phylum)]
    my_genomes = [s * 10 + i for s in my_species for i in range(10)]
    n_local = len(my_genomes)

    words = []
    for g in my_genomes:
        codes = synth.genome_codes(0xA11CE000 + g // 10, G, dev)
        r = synth.splitmix64(((0x5EED0000 + g) << 32) ^ torch.arange(G, dtype=torch.int64, device=dev))
        snp = (synth._lsr(r, 8) % 1000) < 10
        codes = torch.where(snp, (codes + 1 + synth._lsr(r, 40) % 3) % 4, codes)
        words.append(synth.pack_codes_device(codes))
    batch = ops.GenomeBatch.from_device_words(words, [G] * n_local, dev)
    del words, codes, r, snp
    leaves = ctx.alloc_filters(n_local, m)
    torch.cuda.synchronize()

    times = {}

    def timed(name, fn):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) * 1e-3], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        times[name] = float(t.item())
        return out

    ctx.build_filters(batch, k, m, out=leaves)  # warm-up (scratch allocation)
    clocks = ClockSampler(local)
    clocks.__enter__()
    timed("K1_leaf_filters", lambda: ctx.build_filters(batch, k, m, out=leaves))
    rows = [leaves[i] for i in range(n_local)]

    # K2 bottom-up: species, then a chain of levels genus..phylum inside every local phylum (fanout chosen so that
    # 5 levels reduce sp_per_phylum species to 1 phylum filter), then kingdom and db root across ranks
    def unions():
        species = [ctx.bf_or(rows[i * 10:(i + 1) * 10], m) for i in range(len(my_species))]
        level = species
        n_union, n_in = len(species), 10 * len(species)
        per = max(2, round(sp_per_phylum ** (1 / 5)))
        all_levels = [species]
        for depth in range(5):  # genus, family, order, class, phylum
            nxt = []
            groups_per_phylum = len(level) // len(my_phyla)
            for p in range(len(my_phyla)):
                grp = level[p * groups_per_phylum:(p + 1) * groups_per_phylum]
                step = len(grp) if depth == 4 else per
                for i in range(0, len(grp), step):
                    ch = grp[i:i + step]
                    nxt.append(ch[0] if len(ch) == 1 else ctx.bf_or(ch, m))
                    if len(ch) > 1:
                        n_union += 1
                        n_in += len(ch)
            level = nxt
            all_levels.append(level)
        partial = level[0] if len(level) == 1 else ctx.bf_or(level, m)   # this rank's share of the kingdom
        kingdom = shard.or_merge(partial[:ops.filter_words(m)].contiguous(), m, ctx=ctx)
        all_levels.append([kingdom])
        return all_levels, n_union + 1, n_in + len(level)

    unions()
    all_levels, n_union, n_in = timed("K2_unions_bottom_up", unions)

    # K3: all pairs inside every species; species "centroids" (first genome here) among each higher cluster; densities
    def similarities():
        pairs = 0
        for i in range(len(my_species)):
            ctx.bf_allpairs(rows[i * 10:(i + 1) * 10], m)
            pairs += 45
        cents = [rows[i * 10] for i in range(len(my_species))]
        for p in range(len(my_phyla)):
            c = cents[p * sp_per_phylum:(p + 1) * sp_per_phylum][:64]
            if len(c) >= 3:
                ctx.bf_allpairs(c, m)
                pairs += len(c) * (len(c) - 1) // 2
        n_clusters = 0
        for lv in all_levels:
            ctx.bf_popcount(lv, m)
            n_clusters += len(lv)
        return pairs, n_clusters

    similarities()
    pairs, n_clusters = timed("K3_similarities_and_densities", similarities)
    clocks.__exit__(None, None, None)

    tot = torch.tensor([pairs, n_clusters, n_union, n_in, n_local], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(tot)
    if rank == 0:
        pairs, n_clusters, n_union, n_in, n_all = [int(x) for x in tot.tolist()]
        k1_bytes = n_all * ((G + 3) // 4 + fbytes)
        k2_bytes = (n_in + n_union) * fbytes
        out = {
            "workload": "configs[2]: tree build, %d synthetic %.1f Mbp genomes, %d species x 10, k=%d, 2^%d-bit filters, %d GPU(s)"
                        % (n_all, G / 1e6, n_species, k, m.bit_length() - 1, world),
            "n_gpus": world, "genomes": n_all, "clusters": n_clusters, "unions": n_union, "pairs": pairs,
            "seconds": times, "seconds_total": sum(times.values()),
            "genomes_per_s_total": n_all / sum(times.values()),
            "K1": {"Mbp_per_s": n_all * G / times["K1_leaf_filters"] / 1e6,
                   "frac_of_hbm_peak_per_gpu": k1_bytes / times["K1_leaf_filters"] / 1e9 / peak / world},
            "K2": {"algorithmic_GB": k2_bytes / 1e9, "frac_of_hbm_peak_per_gpu": k2_bytes / times["K2_unions_bottom_up"] / 1e9 / peak / world,
                   "note": "includes the cross-GPU bit-range OR-merge of the kingdom filter" if world > 1 else "single GPU: no merge"},
            "K3": {"pairs_per_s": pairs / times["K3_similarities_and_densities"]},
            "clocks": clocks.summary(),
        }
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
