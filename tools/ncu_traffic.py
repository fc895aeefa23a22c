#!/usr/bin/env python
"""profiles/k1_ncu_traffic.json from an `ncu --set full` capture of the K1 kernel:
    tools/ncu_traffic.py gpurun_out/prof.ncu-rep <genomes in the captured launch> [kernel regex]"""
import csv, io, json, os, subprocess, sys
rep, genomes = sys.argv[1], int(sys.argv[2])
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
best = None
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    if len(sys.argv) > 3 and sys.argv[3] not in name:
        continue
    def val(k):
        v, u = float(r[hdr.index(k)]), units[hdr.index(k)]
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
    dur = float(r[hdr.index("gpu__time_duration.sum")])
    if best is None or rd + wr > best[1] + best[2]:
        best = (name, rd, wr, dur, units[hdr.index("gpu__time_duration.sum")])
name, rd, wr, dur, du = best
d = {"kernel": name.split("(")[0], "genomes_in_launch": genomes, "dram_bytes_read": rd, "dram_bytes_write": wr,
     "dram_bytes_per_genome": (rd + wr) / genomes, "duration": "%s %s (under ncu, cold caches)" % (dur, du),
     "source": "ncu --set full: dram__bytes_read.sum + dram__bytes_write.sum, %s" % os.path.basename(rep)}
json.dump(d, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "k1_ncu_traffic.json"), "w"), indent=1)
print(json.dumps(d, indent=1))
