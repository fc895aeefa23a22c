#!/usr/bin/env python
"""cProfile of Database.profile_many on the tools/bench_profile.py database (where does the host time go?)."""
import cProfile, io, os, pstats, sys, tempfile, time, contextlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch
from metasbt_b200 import cli, core, ops, synth
import bench_profile

ctx = ops.default_context(0)
wd = tempfile.mkdtemp(prefix="msbt_prof_")
orig = core.Database.profile_many
state = {}
def wrapped(self, *a, **k):
    if "done" in state:
        return orig(self, *a, **k)
    state["done"] = 1
    pr = cProfile.Profile()
    torch.cuda.synchronize(); t0 = time.perf_counter(); pr.enable()
    out = orig(self, *a, **k)
    torch.cuda.synchronize(); pr.disable(); state["t"] = time.perf_counter() - t0
    s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28); state["stats"] = s.getvalue()
    return out
core.Database.profile_many = wrapped
with contextlib.redirect_stdout(sys.stderr):
    res = bench_profile.run(ctx, workdir=wd)
print("profiled call: %.3f s" % state["t"])
print(state["stats"][:6000])
