#!/usr/bin/env python
"""bench.py -- headline benchmark of the MetaSBT hot path on B200 (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload: Bloom construction (the `howdesbt makebf` step of `metasbt index`,
/root/reference/metasbt/core.py:3920-3931) for --genomes synthetic 5 Mbp genomes per GPU,
k=31, 2^30-bit filters, --min 1.  One step = one pass over the whole per-GPU batch.

* value   : genome Mbp/s, inputs (2-bit packed genomes) already resident in HBM, outputs
            (one 128 MiB filter per genome) left resident in HBM; CUDA events, max over ranks.
* e2e     : the same metric through the public API (pipeline.SketchPipeline, what `metasbt index` drives) with HOST
            buffers, over the WHOLE workload every step: FASTA bytes pinned-host -> device, device ingest, K1, every filter
            back in pinned host memory (dense copy or position lists expanded by host threads; the faster is `e2e`, the
            other `e2e_alt`), with the box's measured device->host and host-write ceilings beside it.
* check   : every rank's first filter, word for word, against the oracle's makebf (a mismatch fails the run).
* mag_queries: BASELINE's second metric (configs[3], tools/bench_query.py) at this N.
* roofline: algorithmic bytes (ceil(G/4) + m/8 per genome) / measured duration
            against MEASURED_PEAKS.json's HBM copy bandwidth.
* cpu_baseline: the oracle's rolling CPU makebf (oracle/, a port -- howdesbt itself is not
            available) on a bounded sample, all host cores, rank 0 at N=1 only.
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "genome Mbp/s into Bloom filters"
UNIT = "Mbp/s"
K = 31
FILTER_BITS = 1 << 30
GENOME_BP = 5_000_000


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--genomes", type=int, default=1000, help="synthetic genomes per GPU (BASELINE configs[1]: 1000)")
    ap.add_argument("--genome-bp", type=int, default=GENOME_BP)
    ap.add_argument("--filter-bits", type=int, default=FILTER_BITS)
    ap.add_argument("--kmer", type=int, default=K)
    ap.add_argument("--variant", type=int, default=0, help="K1 kernel variant (0 auto, 1 direct, 2 tiled, 3 fused, 4 phased)")
    ap.add_argument("--e2e-genomes", type=int, default=0, help="genomes per end-to-end step (0 = the whole workload)")
    ap.add_argument("--e2e-steps", type=int, default=2, help="timed end-to-end passes per transport")
    ap.add_argument("--e2e-chunk", type=int, default=16, help="genomes per pipeline chunk (pinned slot = chunk x filter bytes)")
    ap.add_argument("--e2e-hybrid", action="store_true", help="also time the hybrid transport (measured: within +-25 % of sparse on these boxes)")
    ap.add_argument("--no-numa", action="store_true", help="do not bind the rank to its GPU's NUMA node")
    ap.add_argument("--no-check", action="store_true", help="skip the oracle check of the first filter")
    ap.add_argument("--no-query", action="store_true", help="skip the MAG queries/s section (configs[3])")
    ap.add_argument("--query-leaves", type=int, default=10000)
    ap.add_argument("--query-bits", type=int, default=1 << 25, help="filter size of the query section at every N (fits one GPU)")
    ap.add_argument("--query-bits-full", type=int, default=1 << 28, help="configs[3]'s own filter size, run in addition at N >= 8")
    ap.add_argument("--query-mags", type=int, default=1024, help="MAGs per query step (SURVEY.md 8d, config 4: batches of 1,024)")
    ap.add_argument("--cpu-genomes", type=int, default=0, help="genomes in the CPU-baseline sample (0 = 2 per core)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the K2/K3/K4 measurements (tools/bench_ops.py)")
    return ap.parse_args()


def config_dict(args, n_gpus):
    return {
        "workload": "configs[1]: Bloom construction, %d synthetic %.1f Mbp genomes per GPU, k=%d, 2^%d-bit filters, min=1"
                    % (args.genomes, args.genome_bp / 1e6, args.kmer, args.filter_bits.bit_length() - 1),
        "genomes_per_gpu": args.genomes,
        "genome_bp": args.genome_bp,
        "kmer_size": args.kmer,
        "filter_bits": args.filter_bits,
        "min_kmer_occurrences": 1,
        "sharding": "genomes sharded across %d GPU(s), no data-path collective" % n_gpus,
        "l2": "inputs and outputs (%.1f GiB per GPU) are far larger than L2; no explicit flush"
              % (args.genomes * args.filter_bits / 8 / 2 ** 30),
    }


# ------------------------------------------------------------------ CPU baseline / reference arm


def _cpu_genome_fasta(seed: int, n: int) -> bytes:
    from tests import util
    return util.codes_to_fasta(util.synthetic_codes(seed, n), "g%x" % seed, 80)


def cpu_makebf_throughput(args, n_genomes: int, threads: int):
    """Time the oracle's rolling CPU makebf over `n_genomes` synthetic genomes with `threads` threads.
    ctypes releases the GIL, so a thread pool scales across cores."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import oracle as O
    O.build()
    fastas = [_cpu_genome_fasta(0x5EED0000 + g, args.genome_bp) for g in range(min(n_genomes, threads, 8))]
    work = [fastas[i % len(fastas)] for i in range(n_genomes)]

    def one(fa):
        w = O.makebf(fa, args.kmer, args.filter_bits, 1, rolling=(args.kmer <= 32))
        return int(w[0])

    with ThreadPoolExecutor(max_workers=threads) as ex:
        list(ex.map(one, work[:threads]))  # warm
        t0 = time.perf_counter()
        list(ex.map(one, work))
        dt = time.perf_counter() - t0
    return n_genomes * args.genome_bp / dt / 1e6, dt


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  howdesbt is not available
    (DESIGN.md), so this times the oracle port with all host threads on the same config."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    per_step = args.cpu_genomes or 4 * cores
    vals = []
    for _ in range(args.warmup):
        cpu_makebf_throughput(args, min(per_step, cores), cores)
    t_total = 0.0
    for _ in range(args.steps):
        v, dt = cpu_makebf_throughput(args, per_step, cores)
        vals.append(v)
        t_total += dt
    value = per_step * args.steps * args.genome_bp / t_total / 1e6
    sample = "%d synthetic %.1f Mbp genomes per step (of the %d-genome workload), %d threads" % (
        per_step, args.genome_bp / 1e6, args.genomes, cores)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t_total / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic", "config": config_dict(args, args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "CPU restatement of howdesbt makebf (oracle/msbt_oracle.c: orc_makebf_rolling), in-process, no file I/O; "
                "the real howdesbt binary is not vendored/pinned by the reference and not installed",
    }
    print(json.dumps(line))


# ------------------------------------------------------------------ clocks


class ClockSampler:
    """Samples SM clock and throttle reasons DURING the timed region: NVML (pynvml) every ~10 ms when available,
    else `nvidia-smi` every ~100 ms."""
    QUERY = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index = index
        self.samples = []   # (sm_mhz, max_mhz, [reason names])
        self._stop = threading.Event()
        self._t = None
        self._nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else index
            self._h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self._nvml = pynvml
        except Exception:
            self._nvml = None

    def _sample_nvml(self):
        n = self._nvml
        mhz = n.nvmlDeviceGetClockInfo(self._h, n.NVML_CLOCK_SM)
        mx = n.nvmlDeviceGetMaxClockInfo(self._h, n.NVML_CLOCK_SM)
        r = n.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
        names = []
        for bit, name in ((getattr(n, "nvmlClocksThrottleReasonHwSlowdown", 0x8), "hw_slowdown"),
                          (getattr(n, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40), "hw_thermal_slowdown"),
                          (getattr(n, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20), "sw_thermal_slowdown"),
                          (getattr(n, "nvmlClocksThrottleReasonSwPowerCap", 0x4), "sw_power_cap")):
            if r & bit:
                names.append(name)
        self.samples.append((int(mhz), int(mx), names))

    def _sample_smi(self):
        out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.QUERY,
                              "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
        parts = [p.strip() for p in out.strip().split(",")]
        if len(parts) >= 6 and parts[0].isdigit():
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            self.samples.append((int(parts[0]), int(parts[1]) if parts[1].isdigit() else None,
                                 [nm for i, nm in enumerate(names) if parts[2 + i].lower().startswith("active")]))

    def _run(self):
        while not self._stop.is_set():
            try:
                if self._nvml is not None:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:
                pass
            self._stop.wait(0.01 if self._nvml is not None else 0.1)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=10)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        mhz = sorted(s[0] for s in self.samples)
        reasons = sorted({r for s in self.samples for r in s[2]})
        return {"sm_mhz": mhz[len(mhz) // 2], "sm_max_mhz": self.samples[0][1], "reasons": reasons,
                "samples": len(self.samples), "source": "nvml" if self._nvml is not None else "nvidia-smi"}


# ------------------------------------------------------------------ our arm


def ncu_traffic_per_genome():
    """DRAM bytes (read + write) per genome of the K1 kernel from the committed `ncu --set full` capture
    (profiles/k1_ncu_traffic.json, written by tools/ncu_traffic.py from the .ncu-rep), or None."""
    p = os.path.join(ROOT, "profiles", "k1_ncu_traffic.json")
    try:
        d = json.load(open(p))
        return float(d["dram_bytes_per_genome"]), d.get("source", "")
    except Exception:
        return None, ""


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def host_fasta(args, rank, N, G, device):
    """FASTA text of this rank's N genomes in ONE pinned host buffer (what N FASTA files would hold), generated on the
    device and copied out -- input preparation, not timed.  -> (pinned uint8 tensor, offsets)"""
    import torch
    from metasbt_b200 import synth
    first = synth.fasta_text_device(synth.genome_codes(0x5EED0000 + rank * N, G, device), "g%x" % (0x5EED0000 + rank * N))
    per = first.numel() + 8
    text = torch.empty(per * N, dtype=torch.uint8, pin_memory=True)
    offsets, cur = [0], 0
    for g in range(N):
        seed = 0x5EED0000 + rank * N + g
        t = first if g == 0 else synth.fasta_text_device(synth.genome_codes(seed, G, device), "g%x" % seed)
        text[cur: cur + t.numel()].copy_(t)
        cur += t.numel()
        offsets.append(cur)
    torch.cuda.synchronize()
    return text, offsets


def run_e2e(args, ctx, rank, world, device, barrier, dist, check_filter0):
    """The same metric end to end through the public API (pipeline.SketchPipeline, what `metasbt index` drives):
    FASTA bytes in pinned host memory -> H2D -> device ingest -> K1 -> filters back in pinned host memory, EVERY genome of
    the workload, every step.  Both output transports are measured (dense filters over PCIe; sparse position lists over
    PCIe + host-side rebuild of the dense words); the faster one is `e2e`, the other `e2e_alt`."""
    import numpy as np
    import torch
    from metasbt_b200 import pipeline
    k, m, G, N = args.kmer, args.filter_bits, args.genome_bp, args.genomes
    ne = min(args.e2e_genomes or N, N)
    text, offsets = host_fasta(args, rank, ne, G, device)
    steps = max(1, min(args.steps, args.e2e_steps))
    out = {}
    first_words = {}
    for mode in (("dense", "sparse", "hybrid") if args.e2e_hybrid else ("dense", "sparse")):
        pipe = pipeline.SketchPipeline(ctx, k, m, min_abund=1, chunk_genomes=args.e2e_chunk, slots=2, mode=mode, variant=args.variant)

        def sink(first, words, mode=mode):
            if first == 0 and mode not in first_words:
                first_words[mode] = np.array(words[0], copy=True)
        warm = min(ne, 2 * args.e2e_chunk)
        pipe.run(text, offsets[: warm + 1], sink=sink)   # warm-up: allocates the pinned slots, first-touches them
        pipe.h2d_bytes = pipe.d2h_bytes = 0
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            pipe.run(text, offsets, sink=sink)
        barrier()
        dt = time.perf_counter() - t0
        te = torch.tensor([dt], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        dt = float(te.item())
        out[mode] = {"value": ne * G * world * steps / dt / 1e6, "unit": UNIT, "dense_fraction": round(pipe._dense_fraction, 3) if mode == "hybrid" else None,
                     "h2d_bytes_per_step": int(pipe.h2d_bytes // steps), "d2h_bytes_per_step": int(pipe.d2h_bytes // steps),
                     "genomes_per_step": ne, "steps": steps, "s_per_step": dt / steps, "transport": mode,
                     "pcie_gbs_per_gpu": (pipe.h2d_bytes + pipe.d2h_bytes) / dt / 1e9,
                     "host_filter_bytes_per_step": int(ne * (m // 8))}
        pipe.close()
        del pipe
        torch.cuda.empty_cache()
    # both transports must deliver the very filter the device-resident run built (and the oracle's, checked by the caller)
    for mode, w in first_words.items():
        if not check_filter0(w):
            raise SystemExit("bench.py: e2e (%s) delivered a filter that differs from the device-resident build" % mode)
    # the ceiling this leg runs under: all ranks copying device -> pinned host at once, nothing else going on
    buf_d = torch.empty(1 << 30, dtype=torch.uint8, device=device)
    buf_h = torch.empty(1 << 30, dtype=torch.uint8, pin_memory=True)
    buf_h.copy_(buf_d)
    barrier()
    t0 = time.perf_counter()
    for _ in range(4):
        buf_h.copy_(buf_d, non_blocking=True)
    barrier()
    dt = time.perf_counter() - t0
    tc = torch.tensor([dt], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(tc, op=dist.ReduceOp.MAX)
    d2h_gbs = 4 * (1 << 30) / float(tc.item()) / 1e9
    # ... and the ceiling of the sparse transport: all ranks writing dense filter words into host memory at once (the
    # zero-fill half of msbt_positions_expand on an empty position list, on the threads the pipeline uses), no GPU involved
    from concurrent.futures import ThreadPoolExecutor
    from metasbt_b200 import ops as _ops
    rows = max(1, (1 << 30) // (m // 8))
    host_rows = buf_h.numpy().view(np.uint64)[: rows * (m // 64)].reshape(rows, m // 64)
    empty = np.zeros(0, dtype=np.uint32)
    threads = max(1, min(args.e2e_chunk, os.cpu_count() or 1))
    with ThreadPoolExecutor(max_workers=threads) as ex:
        list(ex.map(lambda i: _ops.positions_expand(empty, m, host_rows[i % rows]), range(threads)))
        barrier()
        t0 = time.perf_counter()
        reps = 4 * max(threads, rows)
        list(ex.map(lambda i: _ops.positions_expand(empty, m, host_rows[i % rows]), range(reps)))
        barrier()
        dt = time.perf_counter() - t0
    tw = torch.tensor([dt], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(tw, op=dist.ReduceOp.MAX)
    host_write_gbs = reps * (m // 8) / float(tw.item()) / 1e9
    best = max(out, key=lambda md: out[md]["value"])
    e2e = out[best]
    e2e["all_transports_mbp_s"] = {md: out[md]["value"] for md in out}
    e2e["note"] = ("FASTA bytes pinned-host->device + device ingest + K1 + every filter back in pinned host memory each step; "
                   "%s transport" % best)
    e2e["d2h_ceiling_gbs_per_gpu_all_ranks_copying"] = d2h_gbs
    e2e["dense_ceiling_mbp_s"] = d2h_gbs * 1e9 / (m // 8) * G * world / 1e6
    e2e["host_filter_write_gbs_per_rank_all_ranks_writing"] = host_write_gbs
    e2e["sparse_ceiling_mbp_s"] = host_write_gbs * 1e9 / (m // 8) * G * world / 1e6
    e2e["ceilings_note"] = ("every genome's result is a %d MiB filter in host memory: the dense transport cannot beat the box's device->host copy rate, "
                            "the sparse one cannot beat the rate at which the host cores write filter words; both are measured here with all ranks "
                            "active and nothing else running" % (m // 8 // 2 ** 20))
    alt = out["dense" if best != "dense" else "sparse"]
    return e2e, alt


def run_ours(args):
    import contextlib
    import numpy as np
    import torch
    import torch.distributed as dist
    from metasbt_b200 import _lib, numa, ops, synth

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    # host threads and pinned buffers on the NUMA node of this rank's GPU (restored before the CPU-baseline leg)
    affinity0 = os.sched_getaffinity(0)
    numa_info = numa.bind_to_gpu_node(local_rank) if not args.no_numa else {"bound": False}
    if world > 1:
        # NCCL may print a version banner on stdout at its first collective: route fd 1 to stderr until then, so that
        # stdout carries the one JSON line only
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=device)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)

    ctx = ops.Context(local_rank)
    k, m, G, N = args.kmer, args.filter_bits, args.genome_bp, args.genomes

    # synthetic inputs, generated and packed on the device (not timed); genome ids are disjoint across ranks
    words = [synth.packed_genome(0x5EED0000 + rank * N + g, G, device) for g in range(N)]
    batch = ops.GenomeBatch.from_device_words(words, [G] * N, device)
    del words
    filters = ctx.alloc_filters(N, m)
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        ctx.build_filters(batch, k, m, out=filters, variant=args.variant)

    for _ in range(args.warmup):
        step()
    barrier()
    launches0 = ctx.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        barrier()
        ev0.record()
        for _ in range(args.steps):
            step()
        ev1.record()
        barrier()
    ms = ev0.elapsed_time(ev1)
    launches = ctx.launch_count - launches0
    t = torch.tensor([ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    total_bp = N * G * world * args.steps
    value = total_bp / (ms_max / 1e3) / 1e6

    # ---- result check inside the bench (a mismatch fails the run): this rank's first filter, bit for bit, against the
    # oracle's makebf of the same genome's FASTA text -- on every rank, so the check also covers the sharded case
    w = ops.filter_words(m)
    got0 = filters[0, :w].cpu().numpy().view(np.uint64).copy()
    check = {"filter0_popcount": int(ctx.bf_popcount([filters[0]], m).cpu()[0])}
    if not args.no_check:
        from oracle import oracle as O   # the checker, outside every timed region
        O.build()
        fa0 = _cpu_genome_fasta(0x5EED0000 + rank * N, G)
        want0 = O.makebf(fa0, k, m, 1, rolling=(k <= 32))
        ok = bool(np.array_equal(got0, want0)) and check["filter0_popcount"] == int(O.bf_popcount(want0))
        flag = torch.tensor([0 if ok else 1], dtype=torch.int64, device=device)
        if world > 1:
            dist.all_reduce(flag)
        if int(flag.item()):
            raise SystemExit("bench.py: the first filter of %d rank(s) differs from the oracle's makebf -- results invalid" % int(flag.item()))
        check["filter0_equals_oracle_makebf"] = True
        check["filter0_word_checksum"] = int(np.bitwise_xor.reduce(want0))
        check["ranks_checked"] = world

    def check_filter0(words0) -> bool:
        return bool(np.array_equal(np.asarray(words0).view(np.uint64)[:w], got0))

    del filters, batch
    torch.cuda.empty_cache()

    # ---- end to end through the public API with host buffers
    e2e = e2e_alt = None
    if not args.no_e2e:
        try:
            e2e, e2e_alt = run_e2e(args, ctx, rank, world, device, barrier, dist, check_filter0)
            if numa_info:
                e2e["numa"] = numa_info
        except Exception as exc:  # (a result MISMATCH raises SystemExit and fails the run; anything else must not cost the headline)
            e2e = {"error": repr(exc)}
            torch.cuda.empty_cache()

    peak, peak_src = measured_peak_gbs()
    # ---- BASELINE's second metric, MAG queries/s, at this N (configs[3])
    mag_queries = None
    if not args.no_query:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import bench_query
        with contextlib.redirect_stdout(sys.stderr):
            try:
                mag_queries = bench_query.run(ctx, rank, world, leaves=args.query_leaves, filter_bits=args.query_bits, mags=args.query_mags,
                                              steps=max(2, min(args.steps, 5)), warmup=1, peak_gbs=peak)
                if world >= 8 and args.query_bits_full and args.query_bits_full != args.query_bits:
                    full = bench_query.run(ctx, rank, world, leaves=args.query_leaves, filter_bits=args.query_bits_full, mags=args.query_mags,
                                           steps=2, warmup=1, peak_gbs=peak)
                    if rank == 0:
                        mag_queries["at_config_filter_size"] = full
            except Exception as exc:  # the headline line must survive a failure here
                mag_queries = {"error": repr(exc)}
        torch.cuda.empty_cache()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    algo_bytes_genome = (G + 3) // 4 + m // 8
    per_genome_us = ms / args.steps / N * 1e3
    achieved = algo_bytes_genome / (per_genome_us * 1e-6) / 1e9
    traffic_genome, traffic_src = ncu_traffic_per_genome()
    launches_per_step = max(1, launches // max(args.steps, 1))
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64", "data": "synthetic", "config": config_dict(args, world),
        "e2e": e2e, "e2e_alt": e2e_alt, "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": (traffic_genome * N if traffic_genome else None), "peak_source": peak_src,
                     "kernel": "kmer_bloom_fused_kernel (msbt_kmer_bloom_build, one launch per %d-genome batch; "
                               "%d launches per step incl. the early-exit redo pass)" % (N, launches_per_step),
                     "algorithmic_bytes_per_launch": algo_bytes_genome * N,
                     "algorithmic_bytes_per_genome": algo_bytes_genome, "us_per_genome": per_genome_us,
                     "traffic_source": traffic_src},
        "clocks": clocks.summary(),
        "lib": os.path.relpath(_lib.loaded_path(), ROOT),
        "check": check,
        "mag_queries": mag_queries,
    }
    if not args.no_secondary and world == 1:
        # the other rows of SURVEY.md 8a (union, popcount distances, MAG queries/s) at 1,024 leaves, 2^28-bit filters,
        # and `metasbt profile` through the Database API (profile_many)
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        with contextlib.redirect_stdout(sys.stderr):
            try:
                import bench_ops
                line["secondary"] = bench_ops.run(ctx, peak_gbs=peak)
            except Exception as exc:  # the headline line must survive a failure here
                line["secondary"] = {"error": repr(exc)}
            try:
                import bench_profile
                line["secondary"]["profile_many"] = bench_profile.run(ctx)
            except Exception as exc:
                line["secondary"]["profile_many"] = {"error": repr(exc)}
    if not args.no_cpu_baseline and world == 1:
        os.sched_setaffinity(0, affinity0)  # the CPU arm gets every host core again
        cores = os.cpu_count() or 1
        n_cpu = args.cpu_genomes or 8 * cores
        v, dt = cpu_makebf_throughput(args, n_cpu, cores)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": "%d of the %d genomes (%.1f Mbp each), %d threads, %.1f s"
                                          % (n_cpu, N, G / 1e6, cores, dt)}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
