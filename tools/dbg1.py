import os, sys, tempfile, json
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from metasbt_b200 import cli, core, backend as B, bffile
from tests import cpu_backend, dataset, snapshot
from oracle import oracle as O
O.build()
wd = tempfile.mkdtemp()
core._BACKENDS.clear()
snap = snapshot.run_config1(wd)
gold = json.load(open("tests/golden/config1.json"))
for k in snap["profiles"]:
    if snap["profiles"][k] != gold["profiles"][k]:
        print("DIFF", k)
        for a, b in zip(snap["profiles"][k], gold["profiles"][k]):
            if a != b: print("   got ", a); print("   want", b)
be = core.get_backend(31, 1 << 22)
ob = cpu_backend.OracleBackend(31, 1 << 22)
mags = [l.strip() for l in open(os.path.join(wd, "inputs", "mags.txt")) if l.strip()]
qb = be.query_positions_many(mags)
oq = ob.query_positions_many(mags)
for i in range(len(mags)):
    p = qb.positions[qb.offsets[i]:qb.offsets[i+1]].cpu().numpy().view(np.uint32).astype(np.uint64)
    print("positions", i, np.array_equal(p, oq[i]), len(p), len(oq[i]))
db = os.path.join(wd, "db")
leaves = sorted(os.path.join(db, "sketches", f) for f in os.listdir(os.path.join(db, "sketches")))
h = be.fetch_hits([be.query_hits_many(qb, list(range(len(mags))), leaves)])[0]
oh = ob.query_hits_many(oq, list(range(len(mags))), leaves)
print("hits equal", np.array_equal(h, oh))
pairs = [(leaves[i], leaves[j]) for i in range(len(leaves)) for j in range(0, len(leaves), 5)]
print("pairs equal", be.dist_pairs(pairs) == ob.dist_pairs(pairs))
for p in leaves:
    ok = np.array_equal(be.filter(p)[:be.words].cpu().numpy().view(np.uint64), bffile.read_bf(p)[1])
    if not ok: print("CACHE != FILE", p)
