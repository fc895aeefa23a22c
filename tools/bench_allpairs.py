#!/usr/bin/env python
"""SURVEY.md 8d stress shape: ONE global all-pairs similarity run over N synthetic genomes (default 10,000 x 5 Mbp,
2^28-bit filters = 320 GB of filters: more than one GPU holds), spread over the GPUs of one box.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 tools/bench_allpairs.py
    python tools/bench_allpairs.py --genomes 4000          # one GPU: what fits (4,000 x 32 MiB = 128 GB)

Every rank builds its block of the leaf filters (K1, untimed), then shard.all_pairs_ring: the local block in one square
all-pairs launch, the other blocks in a half ring of rectangular launches fed by NCCL send/recv -- every pair once.
Time = CUDA events between barriers, max over ranks; clocks sampled during the timed region.  One JSON line on rank 0."""
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
    ap.add_argument("--genomes", type=int, default=10000, help="total over all ranks")
    ap.add_argument("--genome-bp", type=int, default=5_000_000)
    ap.add_argument("--filter-bits", type=int, default=1 << 28)
    ap.add_argument("--chunk", type=int, default=128, help="peer filters per send/recv + rectangular launch")
    ap.add_argument("--route", choices=("bit", "probe"), default="bit",
                    help="'probe' = shard.all_pairs_probe / ops.bf_allpairs_probe (transpose once, every filter's set bits probe the bit-sliced matrix)")
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        keep = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=dev)
        dist.barrier()
        os.dup2(keep, 1)
    ctx = ops.Context(local)
    k, m, G = 31, a.filter_bits, a.genome_bp
    lo, hi = a.genomes * rank // world, a.genomes * (rank + 1) // world
    mat = ctx.alloc_filters(hi - lo, m)
    group = 256
    for g0 in range(lo, hi, group):
        g1 = min(hi, g0 + group)
        words = [synth.pack_codes_device(synth.mutated_codes(0xA11CE000 + g // 10, 0x5EED0000 + g, G, 10, dev)) for g in range(g0, g1)]
        batch = ops.GenomeBatch.from_device_words(words, [G] * (g1 - g0), dev)
        ctx.build_filters(batch, k, m, out=mat[g0 - lo: g1 - lo])
        del words, batch
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # warm-up on a small block (kernel attributes, NCCL channels)
    if a.route == "probe":
        shard.all_pairs_probe(ctx, mat[: min(32, hi - lo)], m, chunk=16)
    else:
        shard.all_pairs_ring(ctx, mat[: min(32, hi - lo)], m, chunk=16)
    barrier()
    launches0 = ctx.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        barrier()
        e0.record()
        if a.route == "probe":
            free, _ = torch.cuda.mem_get_info(dev)
            # one GPU: the matrix in row passes that fit; several: bit-range shards, the sub-filters arrive over send/recv
            full = shard.all_pairs_probe(ctx, mat, m, chunk=a.chunk, slice_bytes=max(1 << 28, free // (2 if world == 1 else 4)))
            out = full[lo:hi]
        else:
            out = shard.all_pairs_ring(ctx, mat, m, chunk=a.chunk)
        e1.record()
        barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    done = torch.tensor([int((out >= 0).sum())], dtype=torch.int64, device=dev)
    # checks: diagonal = popcounts; 10 genomes of a species are far closer to each other than to the next species
    diag = torch.diagonal(out[:, lo:hi])
    pc = ctx.bf_popcount([mat[i] for i in range(min(64, hi - lo))], m)
    ok = bool(torch.equal(diag[: pc.numel()], pc))
    if a.route == "probe":  # cross-check a block against the AND + popcount kernel
        nb = min(48, hi - lo)
        ok = ok and bool(torch.equal(out[:nb, :nb], ctx.bf_allpairs([mat[i] for i in range(nb)], m)))
    # checksum = sum over the entries the ring fills (local blocks both ways, every cross pair once); the probe route fills
    # every cross pair on both sides, so its cross entries count half
    filled = torch.where(out >= 0, out, torch.zeros_like(out))
    local_sum = int(filled[:, lo:hi].sum().item())
    cross = torch.tensor([int(filled.sum().item()) - local_sum], dtype=torch.int64, device=dev)
    csum = torch.tensor([local_sum], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(done)
        dist.all_reduce(csum)
        dist.all_reduce(cross)
    csum = csum + (cross // 2 if a.route == "probe" else cross)
    if rank == 0:
        n = a.genomes
        sec = float(ms.item()) / 1e3
        entries = int(done.item())
        # distinct unordered pairs actually computed: square local blocks compute each pair once but fill both entries
        blocks = [a.genomes * (r + 1) // world - a.genomes * r // world for r in range(world)]
        distinct = sum(b * (b - 1) // 2 for b in blocks) + (n * n - sum(b * b for b in blocks)) // 2
        peak = 6537.6
        try:
            peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
        except Exception:
            pass
        print(json.dumps({
            "workload": "SURVEY 8d stress shape: all pairs among %d synthetic %.1f Mbp genomes, k=31, 2^%d-bit filters (%.0f GB of filters), %d GPU(s)"
                        % (n, G / 1e6, m.bit_length() - 1, n * (m // 8) / 1e9, world),
            "n_gpus": world, "route": a.route, "pairs": distinct, "seconds": sec, "pairs_per_s": distinct / sec,
            "pairs_per_s_per_gpu": distinct / sec / world,
            "read_every_filter_once_frac_of_hbm_peak": n * (m // 8) / sec / 1e9 / peak / world,
            "gpu_launches": ctx.launch_count - launches0, "chunk": a.chunk, "clocks": clocks.summary(),
            "check": {"diagonal_equals_popcounts": ok, "entries_filled": entries, "sum_mod_2_61": int(csum.item() % (1 << 61)),
                      "all_cross_pairs_once": entries == (n * n if a.route == "probe" else
                                                          sum(b * b for b in blocks) + (n * n - sum(b * b for b in blocks)) // 2)},
        }))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
