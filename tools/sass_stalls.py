#!/usr/bin/env python
"""Stall-reason totals over a line range of an ncu source-page CSV: sass_stalls.py file.csv [begin end]"""
import csv, sys
rows=list(csv.reader(open(sys.argv[1])))
h=[i for i,r in enumerate(rows) if 'Source' in r and 'Address' in r][0]
hdr=rows[h]
cols=[c for c in hdr if c.startswith('stall_') and 'Not Issued' not in c]
iE=hdr.index('Instructions Executed')
data=[r for r in rows[h+1:] if len(r)>iE and r[iE].isdigit()]
a=int(sys.argv[2]) if len(sys.argv)>2 else 0
b=int(sys.argv[3]) if len(sys.argv)>3 else len(data)
tot={c:0 for c in cols}
for r in data[a:b]:
    for c in cols:
        v=r[hdr.index(c)]
        tot[c]+=int(v) if v.isdigit() else 0
s=sum(tot.values()) or 1
for c,v in sorted(tot.items(), key=lambda x:-x[1]):
    if v: print(f"{c:28s} {v:8d} {100*v/s:5.1f}%")
