#!/usr/bin/env python
"""BASELINE.json's second metric, "MAG queries/s (1/2/4/8 B200)", on configs[3]: batched profile query of synthetic MAGs
(2 Mbp) against a many-leaf flat SBT whose bit-sliced leaf matrix is sharded by contiguous bit range over the GPUs of one
box.  bench.py runs it at every N (key "mag_queries" of its JSON line); alone:

    python tools/bench_query.py [--leaves 10000 --filter-bits 33554432 --mags 1024 --steps 3 --warmup 1]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29513 tools/bench_query.py ...

Setup (untimed), every rank: builds all leaves (K1; related genomes, 10 per species) and keeps only ITS rows of the
bit-sliced matrix (msbt_bitslice_build on [row_begin, row_end)).
A step = one batch of MAGs: every rank hashes ITS BLOCK of the batch (K1), the distinct positions are exchanged by bit range
(all_to_all), every rank probes its shard with the positions of the whole batch, and the partial hit matrices are summed
with one all_reduce (shard.query_exchange_sharded; `--replicated-hash` gives round 1's scheme: every rank hashes all
MAGs, no exchange).  Time = CUDA events on the launching stream between barriers, max over ranks.  Strong scaling: the
total problem is fixed, the shard shrinks with N.  The default filter size (2^25 bits) is chosen so that the whole
10,000-leaf matrix (43 GB) also fits ONE GPU; at 2^28 bits it needs the 8 GPUs the config names."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run(ctx, rank: int, world: int, leaves: int = 10000, filter_bits: int = 1 << 25, mags: int = 1024,
        genome_bp: int = 5_000_000, mag_bp: int = 2_000_000, steps: int = 3, warmup: int = 1, replicated_hash: bool = False,
        peak_gbs: float = 6537.6, pipelined: bool = True):
    """-> the result dict on rank 0, None elsewhere.  torch.distributed must be initialised when world > 1."""
    import torch
    import torch.distributed as dist
    from metasbt_b200 import ops, shard, synth

    dev = ctx.device
    k, m, L = 31, filter_bits, leaves
    row_begin, row_end = shard.bit_ranges(m, world)[rank]
    lw = ctx.sliced_row_words(L)  # row stride: ceil(L/32) rounded up to 8 words

    # ---- setup (untimed): leaves in groups, only this rank's rows of the sliced matrix are kept
    sliced = torch.empty((row_end - row_begin, lw), dtype=torch.int32, device=dev)
    group = max(32, min(1024, ((24 << 30) // (m // 8)) // 32 * 32))  # multiple of 32: a group fills whole words of every row
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_build = t_slice = 0.0
    for g0 in range(0, L, group):
        g1 = min(L, g0 + group)
        words = [synth.pack_codes_device(synth.mutated_codes(0xA11CE000 + g // 10, 0x5EED0000 + g, genome_bp, 10, dev)) for g in range(g0, g1)]
        batch = ops.GenomeBatch.from_device_words(words, [genome_bp] * (g1 - g0), dev)
        del words
        e0.record()
        lf = ctx.build_filters(batch, k, m)
        e1.record()
        torch.cuda.synchronize()
        t_build += e0.elapsed_time(e1) / 1e3
        e0.record()
        part = ctx.bitslice([lf[i] for i in range(g1 - g0)], row_begin, row_end)
        e1.record()
        torch.cuda.synchronize()
        t_slice += e0.elapsed_time(e1) / 1e3
        sliced[:, g0 // 32:g0 // 32 + part.shape[1]] = part
        del batch, lf, part
    torch.cuda.empty_cache()

    # ---- the MAG batch: 3 % mutants of 2 Mbp prefixes of leaf genomes (source leaf known -> result check)
    src = [(q * 7919 + 13) % L for q in range(mags)]
    qb, qe = (0, mags) if replicated_hash else shard.query_blocks(mags, world)[rank]
    mw = [synth.pack_codes_device(synth.mutated_codes(0xA11CE000 + src[q] // 10, 0xBEEF0000 + q, mag_bp, 30, dev)) for q in range(qb, qe)]
    qbatch = ops.GenomeBatch.from_device_words(mw, [mag_bp] * (qe - qb), dev)
    del mw
    qf = ctx.alloc_filters(max(qe - qb, 1), m)
    qrows = [qf[i] for i in range(qe - qb)]
    timings = dict()

    def step(tm=None):
        if qe > qb:
            ctx.build_filters(qbatch, k, m, out=qf)  # K1 on this rank's block of the batch
        if replicated_hash:
            return shard.query_sharded(ctx, sliced, L, m, qrows, cap=mags * mag_bp)
        return shard.query_exchange_sharded(ctx, sliced, L, m, qrows, timings=tm, cap=(qe - qb) * mag_bp, batch_size=mags)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # Batches are pipelined over two streams: while the shard is probed with batch i (main stream), batch i + 1 is hashed,
    # its positions extracted and exchanged (side stream) -- what a stream of query batches looks like in production.
    # A context is driven from one stream at a time (its scratch is shared by its calls): the side stream gets its own.
    main, side = torch.cuda.current_stream(dev), torch.cuda.Stream(dev, priority=-1 if os.environ.get("MSBT_QPIPE_PRIO", "1") != "0" else 0)
    ctx_side = ops.Context(dev.index) if pipelined and not replicated_hash else ctx

    def prepare(first=False):
        if first:
            side.wait_stream(main)  # (later batches must NOT wait for the probe that was just queued on the main stream)
        with torch.cuda.stream(side):
            if qe > qb:
                ctx_side.build_filters(qbatch, k, m, out=qf)
            return shard.query_exchange_prepare(ctx_side, m, qrows, cap=(qe - qb) * mag_bp, batch_size=mags)

    def run_steps(n):
        out = None
        ex = prepare(first=True)
        for i in range(n):
            main.wait_stream(side)
            out = shard.query_exchange_finish(ctx, sliced, L, m, ex, reduce=False)   # probe, asynchronous
            if i + 1 < n:
                ex = prepare()                                                        # next batch under the probe
            shard.all_reduce_counts(out)
        return out

    for _ in range(max(warmup, 1)):
        step()
    if not replicated_hash and pipelined:
        run_steps(2)
    barrier()
    launches0 = ctx.launch_count + (ctx_side.launch_count if ctx_side is not ctx else 0)
    e0.record()
    if not replicated_hash and pipelined:
        hits = run_steps(steps)
    else:
        for _ in range(steps):
            hits = step()
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    sec = float(ms.item()) / 1e3
    launches = ctx.launch_count + (ctx_side.launch_count if ctx_side is not ctx else 0) - launches0
    # one more step with per-phase events (synchronises between phases, so it is not part of the timed region)
    if not replicated_hash:
        step(timings)
    e0.record()
    if qe > qb:
        ctx.build_filters(qbatch, k, m, out=qf)
    e1.record()
    torch.cuda.synchronize()
    timings["ms_hash_local_block"] = e0.elapsed_time(e1)

    best = hits.argmax(dim=1).cpu().tolist()
    same_species = sum(1 for q, g in enumerate(src) if best[q] // 10 == g // 10)
    checksum = int((hits.to(torch.int64) * (torch.arange(L, device=dev, dtype=torch.int64) % 1009 + 1)).sum().item())
    n_pos_local = torch.tensor([int(sum(int(c) for c in ctx.bf_popcount(qrows, m).tolist())) if qrows else 0], dtype=torch.int64, device=dev)
    if world > 1 and not replicated_hash:
        dist.all_reduce(n_pos_local)
    n_pos = int(n_pos_local.item())
    if rank != 0:
        return None
    in_range = int(timings.get("positions_in_shard", 0))
    t_probe = timings.get("ms_probe", 0.0) / 1e3
    algo = in_range * ((L + 7) // 8) + mags * 4 * L  # one row gather per distinct position + the hit vector out
    roof = None
    if t_probe > 0:
        roof = {"bound": "hbm", "kernel": "query_sliced_kernel (rank 0's shard, one launch per step)", "achieved": algo / t_probe / 1e9,
                "peak": peak_gbs, "unit": "GB/s", "frac": algo / t_probe / 1e9 / peak_gbs, "ms_probe": t_probe * 1e3,
                "algorithmic_bytes": algo, "positions_total": n_pos, "positions_in_shard": in_range}
    return {
        "metric": "MAG queries/s (batched profile query, bit-sliced flat SBT sharded by bit range)",
        "value": mags * steps / sec, "unit": "MAG/s", "n_gpus": world, "steps": steps, "warmup": max(warmup, 1),
        "ms_per_step": sec * 1e3 / steps, "higher_is_better": True, "scaling": "strong", "dtype": "u32", "data": "synthetic",
        "config": {"workload": "configs[3]: %d MAGs x %.1f Mbp per step against %d leaves (5 Mbp genomes, 10 per species), k=31, 2^%d-bit filters"
                               % (mags, mag_bp / 1e6, L, m.bit_length() - 1),
                   "sharding": ("rows [r*m/N, (r+1)*m/N) of the bit-sliced matrix per rank; " +
                                ("hashing replicated on every rank, extraction and probe per bit range" if replicated_hash else
                                 "MAG hashing sharded across ranks, positions exchanged by bit range (all_to_all), probe per bit range") +
                                "; hits all_reduce(SUM)" + ("; batches pipelined over two streams (hash + exchange of batch i+1 under the probe of batch i)"
                                                            if pipelined and not replicated_hash else "")),
                   "l2": "the shard (%.1f GB) is far larger than L2" % (sliced.numel() * 4 / 1e9)},
        "gpu_launches": launches,
        "roofline": roof,
        "phases_ms_one_step": {k_: round(v, 3) for k_, v in timings.items() if k_.startswith("ms_")},
        "setup": {"s_build_leaves": t_build, "s_bitslice": t_slice, "sliced_GB_per_rank": sliced.numel() * 4 / 1e9},
        "check": {"best_hit_in_source_species": "%d/%d" % (same_species, mags), "hits_checksum": checksum}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--leaves", type=int, default=10000)
    ap.add_argument("--filter-bits", type=int, default=1 << 25)
    ap.add_argument("--mags", type=int, default=1024)
    ap.add_argument("--genome-bp", type=int, default=5_000_000)
    ap.add_argument("--mag-bp", type=int, default=2_000_000)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--replicated-hash", action="store_true")
    ap.add_argument("--no-pipeline", action="store_true", help="one batch after the other on one stream")
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    from metasbt_b200 import ops
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
    peak = 6537.6
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    res = run(ops.Context(local), rank, world, a.leaves, a.filter_bits, a.mags, a.genome_bp, a.mag_bp, a.steps, a.warmup,
              a.replicated_hash, peak, not a.no_pipeline)
    if rank == 0:
        print(json.dumps(res))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
