#!/usr/bin/env python
"""Opcode / hot-line histogram of an `ncu --page source --csv --print-source sass` export."""
import csv, sys, collections
rows=list(csv.reader(open(sys.argv[1])))
top=int(sys.argv[2]) if len(sys.argv)>2 else 25
hi=[i for i,r in enumerate(rows) if 'Source' in r and 'Address' in r]
for h in hi[:1]:
    hdr=rows[h]
    iS=hdr.index('Source'); iN=hdr.index('# Samples'); iE=hdr.index('Instructions Executed')
    iW=hdr.index('L1 Wavefronts Shared'); iWI=hdr.index('L1 Wavefronts Shared Ideal')
    data=[]
    for r in rows[h+1:]:
        if len(r)<=iE or not r[iE].isdigit(): continue
        data.append((r[iS].strip(), int(r[iN] or 0), int(r[iE] or 0), int(r[iW] or 0), int(r[iWI] or 0)))
    tot_s=sum(d[1] for d in data) or 1; tot_e=sum(d[2] for d in data) or 1
    print("instr lines",len(data),"samples",tot_s,"inst exec",tot_e)
    ce=collections.Counter(); cs=collections.Counter()
    for s,n,e,w,wi in data:
        t=s.split()
        op=t[1] if t[0].startswith('@') else t[0]
        op=op.split('.')[0]
        ce[op]+=e; cs[op]+=n
    for op,e in ce.most_common(top): print(f"{op:10s} exec {e:9d} {100*e/tot_e:5.1f}%  samples {cs[op]:6d} {100*cs[op]/tot_s:5.1f}%")
    print("shared wavefronts", sum(d[3] for d in data), "ideal", sum(d[4] for d in data))
    print("--- top sampled lines")
    for i,(s,n,e,w,wi) in sorted(enumerate(data), key=lambda x:-x[1][1])[:top]:
        print(f"{i:5d} samples {n:5d} exec {e:8d} shw {w:8d}/{wi:8d}  {s[:90]}")
