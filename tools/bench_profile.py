#!/usr/bin/env python
"""`metasbt profile` through the product API (Database.profile_many, the batch form of the process pool of
metasbt.py:1124-1143) on a synthetic 7-level database: wall-clock MAGs/s from FASTA files on disk to profile tables on
disk, against the per-genome Database.profile on a sample of the same MAGs.  bench.py reports it as
secondary.profile_many (N = 1); alone:

    python tools/bench_profile.py [--species 64 --genomes-per-species 4 --genome-bp 50000 --filter-bits 33554432 --mags 512]
"""
import argparse
import json
import os
import shutil
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _taxonomy(s: int) -> str:
    # a binary tree above the species: 1 kingdom, 2 phyla, 4 classes, 8 orders, 16 families, 32 genera
    return "|".join(["k__Synth", "p__P%d" % (s % 2), "c__C%d" % (s % 4), "o__O%d" % (s % 8), "f__F%d" % (s % 16),
                     "g__G%d" % (s % 32), "s__S%d" % s])


def run(ctx, species: int = 64, per_species: int = 4, genome_bp: int = 50_000, filter_bits: int = 1 << 25, mags: int = 512,
        mag_bp: int = 25_000, single_sample: int = 32, workdir: str = None):
    """Defaults: 12.8 M k-mers in the database against 2^25-bit filters -- the root filter is about a third full, the regime
    MetaSBT's own filter-size estimate (the number of distinct k-mers, core.py:470-529) puts a database in; with a filter
    that saturates at the upper levels every cluster matches every query and the descent visits the whole tree."""
    import torch
    from metasbt_b200 import cli, core, synth
    from metasbt_b200 import ops
    dev = ctx.device
    ctx = ops.default_context(dev.index)  # the context the Database's backend launches on (bench.py holds its own)
    own = workdir is None
    workdir = workdir or tempfile.mkdtemp(prefix="msbt_profile_bench_")
    try:
        gdir = os.path.join(workdir, "genomes")
        os.makedirs(gdir, exist_ok=True)
        refs = os.path.join(workdir, "references.tsv")
        with open(refs, "w") as tsv:
            for s in range(species):
                for i in range(per_species):
                    g = s * per_species + i
                    codes = synth.mutated_codes(0xA11CE000 + s, 0x5EED0000 + g, genome_bp, 10, dev)
                    p = os.path.join(gdir, "REF_%05d.fna" % g)
                    with open(p, "wb") as f:
                        f.write(synth.fasta_text_device(codes, "REF_%05d" % g).cpu().numpy().tobytes())
                    tsv.write("%s\t%s\n" % (p, _taxonomy(s)))
        mag_paths = []
        for q in range(mags):
            s = (q * 37) % species
            codes = synth.mutated_codes(0xA11CE000 + s, 0xBEEF0000 + q, mag_bp, 30, dev)
            p = os.path.join(gdir, "MAG_%05d.fna" % q)
            with open(p, "wb") as f:
                f.write(synth.fasta_text_device(codes, "MAG_%05d" % q).cpu().numpy().tobytes())
            mag_paths.append(p)
        core._BACKENDS.clear()
        t0 = time.perf_counter()
        app = cli.MetaSBT(["index", "--workdir", workdir, "--database", "db", "--references", refs, "--kmer-size", "31",
                           "--filter-size", str(filter_bits), "--min-kmer-occurrences", "1", "--nproc", "1"])
        t_index = time.perf_counter() - t0
        db = app.database
        entries = [core.Entry(db, os.path.splitext(os.path.basename(p))[0], os.path.splitext(os.path.basename(p))[0], "genome") for p in mag_paths]
        t0 = time.perf_counter()
        sketches = core.Entry.sketch_many(entries, mag_paths)
        t_sketch = time.perf_counter() - t0
        prof_dir = os.path.join(workdir, "tmp", "profiles")
        launches0 = ctx.launch_count
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        tables = db.profile_many(mag_paths, sketches, uncertainty=20.0)
        torch.cuda.synchronize()
        t_many = time.perf_counter() - t0
        launches_many = ctx.launch_count - launches0
        # the per-genome method on a sample, from scratch
        shutil.rmtree(prof_dir)
        ns = min(single_sample, mags)
        launches0 = ctx.launch_count
        t0 = time.perf_counter()
        single = [db.profile(p, s, uncertainty=20.0) for p, s in zip(mag_paths[:ns], sketches[:ns])]
        torch.cuda.synchronize()
        t_single = time.perf_counter() - t0
        launches_single = ctx.launch_count - launches0
        right = sum(1 for q, t in enumerate(tables) if list(t["species"].keys()) and
                    min(t["species"], key=t["species"].get).endswith("|s__S%d" % ((q * 37) % species)))
        return {
            "config": {"species": species, "genomes": species * per_species, "genome_bp": genome_bp, "filter_bits": filter_bits,
                       "mags": mags, "mag_bp": mag_bp, "levels": 7, "kmer_size": 31},
            "profile_many_mags_per_s": mags / t_many, "profile_many_s": t_many, "profile_many_launches": launches_many,
            "per_genome_profile_mags_per_s": ns / t_single, "per_genome_launches_per_mag": launches_single / ns,
            "batch_speedup": (mags / t_many) / (ns / t_single),
            "index_s": t_index, "sketch_mags_s": t_sketch,
            "trees_probed_per_mag": sum(len(t[level]) for t in tables for level in t) / max(len(tables), 1),
            "check": {"batch_equals_per_genome": single == tables[:ns], "closest_species_is_source": "%d/%d" % (right, mags)},
            "note": "wall clock through Database.profile_many: FASTA files on disk -> ingest + K1 + extraction for the whole batch, "
                    "8 level-synchronous tree probes (one device->host copy each), one pair-list distance launch, tables on disk"}
    finally:
        if own:
            shutil.rmtree(workdir, ignore_errors=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--species", type=int, default=64)
    ap.add_argument("--genomes-per-species", type=int, default=4)
    ap.add_argument("--genome-bp", type=int, default=50_000)
    ap.add_argument("--filter-bits", type=int, default=1 << 25)
    ap.add_argument("--mags", type=int, default=512)
    a = ap.parse_args()
    from metasbt_b200 import ops
    print(json.dumps(run(ops.default_context(0), a.species, a.genomes_per_species, a.genome_bp, a.filter_bits, a.mags, a.genome_bp // 2)))
