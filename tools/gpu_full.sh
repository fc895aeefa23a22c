#!/bin/bash
# full measurement pass on the GPU box (one GPU): GPU test suite, default bench, reference arm, ncu launch list of the bench,
# ncu --set full of the K1 kernel and of the sliced probe.   usage: tools/gpu_full.sh <tag>
tag=${1:-r02}
mkdir -p gpurun_out
timeout -k 5 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?"; grep -v "^Update" gpurun_out/pytest_$tag.log | tail -3
timeout -k 5 900 python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/bench_$tag.json
timeout -k 5 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_${tag}_reference.json 2> gpurun_out/bench_${tag}_reference.err; echo "reference rc=$?"; cut -c1-300 gpurun_out/bench_${tag}_reference.json
# launch list of the same bench (bounded: 128 genomes per e2e pass, one e2e pass, 64 MAGs per query step): gpu__time_duration only, no replay
timeout -k 5 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"kmer_|bf_|query_|bitslice|overflow|tile_build|fasta_|fp_" -c 4000 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-secondary --e2e-genomes 128 --e2e-steps 1 --query-mags 64 > gpurun_out/ncu_launches_$tag.log 2>&1; echo "ncu launches rc=$?"
timeout -k 5 300 ncu --set full --clock-control none --import-source on -k regex:kmer_bloom_fused -c 1 -o gpurun_out/prof_${tag}_k1 \
    python bench.py --genomes 48 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --no-secondary --no-query --no-check > gpurun_out/ncu_full_$tag.log 2>&1; echo "ncu K1 rc=$?"
timeout -k 5 300 ncu --set full --clock-control none --import-source on -k regex:kmer_bloom_fused -c 1 -o gpurun_out/prof_${tag}_k1_2p28 \
    python bench.py --genomes 48 --filter-bits 268435456 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --no-secondary --no-query --no-check > gpurun_out/ncu_full_${tag}_2p28.log 2>&1; echo "ncu K1 2^28 rc=$?"
