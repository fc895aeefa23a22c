#!/usr/bin/env python
"""BASELINE.json configs[3]: batched profile query of synthetic MAGs (2 Mbp) against a many-leaf flat SBT whose bit-sliced
leaf matrix is sharded by contiguous bit range over the GPUs of one box.

    python tools/bench_query.py [--leaves 10000 --filter-bits 33554432 --mags 256 --steps 3 --warmup 1]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29513 tools/bench_query.py ...

Every rank: builds all leaves (K1; related genomes, 10 per species), keeps only ITS rows of the bit-sliced matrix
(msbt_bitslice_build on [row_begin, row_end)), hashes all MAGs of the batch (K1, redundant across ranks as SURVEY 8e
proposes), extracts the positions of its own bit range and probes its shard (shard.query_sharded); the partial hit
matrices are summed with one NCCL all_reduce.  A step = one batch of MAGs through hash -> extract -> probe -> all_reduce; time = CUDA events on the
launching stream between barriers, max over ranks.  Strong scaling: the total problem is fixed, the shard shrinks with N.
The default filter size (2^25 bits) is chosen so that the whole 10,000-leaf matrix (42 GB) also fits ONE GPU; at 2^28 bits
it needs the 8 GPUs the config names.  Prints one JSON line on rank 0."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

from metasbt_b200 import ops, shard, synth


def mutated(seed_anc: int, seed_mut: int, n: int, per_mille: int, dev) -> torch.Tensor:
    codes = synth.genome_codes(seed_anc, n, dev)
    r = synth.splitmix64((seed_mut << 32) ^ torch.arange(n, dtype=torch.int64, device=dev))
    snp = (synth._lsr(r, 8) % 1000) < per_mille
    return synth.pack_codes_device(torch.where(snp, (codes + 1 + synth._lsr(r, 40) % 3) % 4, codes))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--leaves", type=int, default=10000)
    ap.add_argument("--filter-bits", type=int, default=1 << 25)
    ap.add_argument("--mags", type=int, default=256)
    ap.add_argument("--genome-bp", type=int, default=5_000_000)
    ap.add_argument("--mag-bp", type=int, default=2_000_000)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        keep = os.dup(1)
        os.dup2(2, 1)  # NCCL's banner goes to stderr, the JSON line stays alone on stdout
        dist.init_process_group("nccl", device_id=dev)
        dist.barrier()
        os.dup2(keep, 1)
    ctx = ops.Context(local)
    k, m, L = 31, a.filter_bits, a.leaves
    row_begin, row_end = shard.bit_ranges(m, world)[rank]
    lw = ctx.sliced_row_words(L)  # row stride: ceil(L/32) rounded up to 8 words

    # ---- setup (untimed): leaves in groups, only this rank's rows of the sliced matrix are kept
    sliced = torch.empty((row_end - row_begin, lw), dtype=torch.int32, device=dev)
    group = 1024  # multiple of 32: a group fills whole words of every row
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_build = t_slice = 0.0
    for g0 in range(0, L, group):
        g1 = min(L, g0 + group)
        words = [mutated(0xA11CE000 + g // 10, 0x5EED0000 + g, a.genome_bp, 10, dev) for g in range(g0, g1)]
        batch = ops.GenomeBatch.from_device_words(words, [a.genome_bp] * (g1 - g0), dev)
        del words
        e0.record()
        leaves = ctx.build_filters(batch, k, m)
        e1.record()
        torch.cuda.synchronize()
        t_build += e0.elapsed_time(e1) / 1e3
        e0.record()
        part = ctx.bitslice([leaves[i] for i in range(g1 - g0)], row_begin, row_end)
        e1.record()
        torch.cuda.synchronize()
        t_slice += e0.elapsed_time(e1) / 1e3
        sliced[:, g0 // 32:g0 // 32 + part.shape[1]] = part
        del batch, leaves, part
    torch.cuda.empty_cache()

    # ---- the MAG batch: 3 % mutants of 2 Mbp prefixes of leaf genomes (source leaf known -> result check)
    src = [(q * 7919 + 13) % L for q in range(a.mags)]
    mw = [mutated(0xA11CE000 + g // 10, 0xBEEF0000 + q, a.mag_bp, 30, dev) for q, g in enumerate(src)]
    qbatch = ops.GenomeBatch.from_device_words(mw, [a.mag_bp] * a.mags, dev)
    del mw
    qf = ctx.alloc_filters(a.mags, m)
    launches0 = None

    qrows = [qf[i] for i in range(a.mags)]

    def step():
        ctx.build_filters(qbatch, k, m, out=qf)  # every rank hashes the whole batch (K1)
        return shard.query_sharded(ctx, sliced, L, m, qrows, cap=a.mags * a.mag_bp)  # extract own range, probe, all_reduce

    for _ in range(a.warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches0 = ctx.launch_count
    e0.record()
    for _ in range(a.steps):
        hits = step()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    sec = float(ms.item()) / 1e3
    launches = ctx.launch_count - launches0

    # probe-only time of this rank's shard (the dominant kernel) for the roofline
    w0, w1 = row_begin // 64, (row_end + 63) // 64
    pos, offs = ctx.extract_positions_packed([f[w0:w1] for f in qrows], (w1 - w0) * 64, cap=a.mags * a.mag_bp)
    e0.record()
    ctx.query_sliced_packed(sliced, L, 0, row_end - row_begin, pos, offs)
    e1.record()
    torch.cuda.synchronize()
    t_probe = e0.elapsed_time(e1) / 1e3
    in_range = int(offs[-1])
    n_pos = int(sum(int(c) for c in ctx.bf_popcount(qrows, m).tolist()))
    best = hits.argmax(dim=1).cpu().tolist()
    same_species = sum(1 for q, g in enumerate(src) if best[q] // 10 == g // 10)
    checksum = int((hits.to(torch.int64) * (torch.arange(L, device=dev, dtype=torch.int64) % 1009 + 1)).sum().item())
    if rank == 0:
        peak = 6537.6
        try:
            peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
        except Exception:
            pass
        algo = in_range * ((L + 7) // 8) + a.mags * 4 * L  # one row gather per distinct position + the hit vector out
        print(json.dumps({
            "metric": "MAG queries/s (batched profile query, bit-sliced flat SBT sharded by bit range)",
            "value": a.mags * a.steps / sec, "unit": "MAG/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": sec * 1e3 / a.steps, "higher_is_better": True, "scaling": "strong", "dtype": "u32", "data": "synthetic",
            "config": {"workload": "configs[3]: %d MAGs x %.1f Mbp per step against %d leaves (5 Mbp genomes, 10 per species), k=31, 2^%d-bit filters"
                                   % (a.mags, a.mag_bp / 1e6, L, m.bit_length() - 1),
                       "sharding": "rows [r*m/N, (r+1)*m/N) of the bit-sliced matrix per rank; hashing replicated, extraction and probe per bit range; hits all_reduce(SUM)",
                       "l2": "the shard (%.1f GB) is far larger than L2" % (sliced.numel() * 4 / 1e9)},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": "query_sliced_kernel (rank 0's shard)", "achieved": algo / t_probe / 1e9, "peak": peak,
                         "unit": "GB/s", "frac": algo / t_probe / 1e9 / peak, "ms_probe": t_probe * 1e3,
                         "algorithmic_bytes": algo, "positions_total": n_pos, "positions_in_shard": in_range},
            "setup": {"s_build_leaves": t_build, "s_bitslice": t_slice, "sliced_GB_per_rank": sliced.numel() * 4 / 1e9},
            "check": {"best_hit_in_source_species": "%d/%d" % (same_species, a.mags), "hits_checksum": checksum}}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
