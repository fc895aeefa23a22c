#!/bin/bash
# usage: tools/gpu_qlibs.sh <mags> lib1.so lib2.so ...   configs[3] query bench with alternative kernel builds
n=$1; shift
for lib in "$@"; do
  MSBT_LIB=$PWD/$lib timeout 200 python tools/bench_query.py --mags $n --steps 2 --warmup 1 > gpurun_out/ql.json 2> gpurun_out/ql.err || { echo "$lib FAILED"; tail -3 gpurun_out/ql.err; continue; }
  python - "$lib" <<'PY'
import json,sys
d=json.loads(open("gpurun_out/ql.json").read().strip().splitlines()[-1])
print("%-36s %.0f MAG/s  step %.1f ms  probe %.1f ms  checksum %d" % (sys.argv[1], d["value"], d["ms_per_step"], d["phases_ms_one_step"]["ms_probe"], d["check"]["hits_checksum"]))
PY
done
