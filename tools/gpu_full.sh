#!/bin/bash
# full measurement pass on the GPU box: default bench, reference arm, ncu launch list, ncu --set full of the K1 kernel
tag=${1:-r01}
mkdir -p gpurun_out
timeout -k 5 600 python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/bench_$tag.json
timeout -k 5 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_${tag}_reference.json 2> gpurun_out/bench_${tag}_reference.err; echo "reference rc=$?"; cut -c1-300 gpurun_out/bench_${tag}_reference.json
timeout -k 5 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"kmer_|bf_|query_|bitslice|overflow|tile_build" -c 300 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --genomes 64 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches_$tag.log 2>&1; echo "ncu launches rc=$?"
timeout -k 5 300 ncu --set full --clock-control none --import-source on -k regex:kmer_bloom_fused -c 1 -o gpurun_out/prof_${tag}_k1 python bench.py --genomes 48 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --no-secondary > gpurun_out/ncu_full_$tag.log 2>&1; echo "ncu full rc=$?"
