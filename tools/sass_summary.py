#!/usr/bin/env python
"""Per-kernel SASS summary of the shipped library (no GPU needed):

    python tools/sass_summary.py [metasbt_b200/libmsbt.so] > profiles/rNN_sass_summary.txt

For every sm_100a kernel: instruction count, registers / shared memory / stack from the resource usage dump, and the
counts of the mnemonics that tell how it moves data and what bounds it (bulk-tensor / bulk copies of the TMA engine,
cp.async, reductions and atomics, POPC / LOP3 / IMAD, shuffles, barriers)."""
import collections
import re
import subprocess
import sys

so = sys.argv[1] if len(sys.argv) > 1 else "metasbt_b200/libmsbt.so"
KEYS = ["UBLKCP", "UTMA", "LDGSTS", "LDG", "STG", "LDS", "STS", "REDG", "ATOMG", "ATOMS", "FENCE", "POPC", "LOP3", "IMAD", "SHF", "SHFL",
        "BAR", "MEMBAR", "NANOSLEEP", "BRA"]
res = subprocess.run(["cuobjdump", "--dump-resource-usage", so], capture_output=True, text=True).stdout
usage = {}
name = None
for line in res.splitlines():
    m = re.search(r"Function (\S+):", line)
    if m:
        name = m.group(1)
    elif name and "REG:" in line:
        usage[name] = " ".join(t for t in line.split() if t.split(":")[0] in ("REG", "STACK", "SHARED"))
        name = None
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
demangle = {}
names = sorted(set(re.findall(r"Function : (\S+)", sass)))
try:
    out = subprocess.run(["cu++filt"] + names, capture_output=True, text=True).stdout.splitlines()
    demangle = dict(zip(names, out))
except Exception:
    pass
cur, counts, total = None, {}, {}
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        total[cur] = 0
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P[0-9T]+\s+)?([A-Z0-9_.]+)", line)
    if cur and m:
        op = m.group(1).split(".")[0]
        total[cur] += 1
        counts[cur][op] += 1
print("SASS summary of %s (cuobjdump -sass, sm_100a)\n" % so)
for k in sorted(counts, key=lambda x: demangle.get(x, x)):
    c = counts[k]
    full = demangle.get(k, k)
    short = full[: full.rindex(">(") + 1] if ">(" in full else re.sub(r"\(.*", "", full)
    print("%s\n    %d instructions; %s" % (short, total[k], usage.get(k, "")))
    print("    " + "  ".join("%s %d" % (op, sum(v for o, v in c.items() if o == op or (op in ("UTMA",) and o.startswith(op))))
                            for op in KEYS if any(o == op or (op == "UTMA" and o.startswith(op)) for o in c)))
tot = collections.Counter()
for c in counts.values():
    tot.update(c)
print("\nwhole library: %d kernels, %d instructions; UBLKCP %d (TMA-engine bulk copies), UTMACMDFLUSH %d, LDGSTS %d (cp.async), POPC %d, REDG %d, ATOMG %d, ATOMS %d"
      % (len(counts), sum(total.values()), tot["UBLKCP"], tot["UTMACMDFLUSH"], tot["LDGSTS"], tot["POPC"], tot["REDG"], tot["ATOMG"], tot["ATOMS"]))
