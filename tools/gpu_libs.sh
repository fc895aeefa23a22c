#!/bin/bash
# usage: tools/gpu_libs.sh <genomes> lib1.so lib2.so ...   bench variant 3 with alternative kernel builds
n=$1; shift
for lib in "$@"; do
  MSBT_LIB=$PWD/$lib timeout 60 python bench.py --genomes $n --steps 3 --warmup 2 --variant 3 --no-e2e --no-cpu-baseline > gpurun_out/lib.log 2> gpurun_out/lib.err || { echo "$lib FAILED"; tail -5 gpurun_out/lib.err; continue; }
  python - "$lib" <<'PY'
import json,sys
d=json.loads(open("gpurun_out/lib.log").read().strip().splitlines()[-1])
print("%-40s %.2f us/genome  frac %.3f popc %d" % (sys.argv[1], d["roofline"]["us_per_genome"], d["roofline"]["frac"], d["check"]["filter0_popcount"]))
PY
done
