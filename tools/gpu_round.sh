#!/bin/bash
# usage: tools/gpu_round.sh <tag>   (runs on the GPU box from the repo root)
tag=${1:-x}
mkdir -p gpurun_out
timeout -k 5 ${PYTEST_TIMEOUT:-150} python -m pytest tests -m gpu -x -q ${PYTEST_ARGS:-} > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_$tag.log
for v in ${VARIANTS:-4 3}; do
  timeout -k 5 60 python bench.py --genomes 200 --steps 3 --warmup 3 --variant $v --no-e2e --no-cpu-baseline > gpurun_out/bench_${tag}_v$v.log 2> gpurun_out/bench_${tag}_v$v.err; echo "bench v$v rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_${tag}_v$v.log").read().strip().splitlines()[-1])
    print("variant $v: %.1f Mbp/s  %.2f us/genome  frac %.3f launches %d popc %d" % (d["value"], d["roofline"]["us_per_genome"], d["roofline"]["frac"], d["gpu_launches"], d["check"]["filter0_popcount"]))
except Exception as e:
    print("no result", e); print(open("gpurun_out/bench_${tag}_v$v.err").read()[-800:])
PY
done
