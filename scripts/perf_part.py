"""Partitioned-table kernels on one GPU (n_part partitions, no exchange): per-kernel times.
usage: perf_part.py [cfg] [k] [n_part] [prefilter 0|1]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from superplus_b200 import api, synth
from superplus_b200 import dist as gdist
name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
k = int(sys.argv[2]) if len(sys.argv) > 2 else 25
n_part = int(sys.argv[3]) if len(sys.argv) > 3 else 2
pre = bool(int(sys.argv[4])) if len(sys.argv) > 4 else True
inp = synth.make_config(name)
ctx = api.Context(0)
cs, rs = ctx.upload(inp.contigs), ctx.upload(inp.reads)
# one table holding everything plays all owners: the routing does not care who answers
t = ctx.table_build(cs, k)
flt = None
if pre:
    nw, k3 = ctx.filter_shape(cs.kmers(k))
    ft = torch.zeros(nw, dtype=torch.int32, device="cuda")
    t.filter_add(ft.data_ptr(), nw, k3)
    flt = (ft.data_ptr(), nw, k3)
ctx.prof(True)
for it in range(3):
    ctx.prof_reset()
    r = ctx.route_plan(rs, k, n_part, 0, rs.tiles, prefilter=flt)
    n = r.kmers
    keys = torch.empty(max(n, 1), dtype=torch.int64, device="cuda"); ans = torch.empty(max(n, 1), dtype=torch.int64, device="cuda")
    r.keys(keys.data_ptr())
    t.lookup_keys(keys.data_ptr(), n, ans.data_ptr())
    h = r.collect(ans.data_ptr())
    ctx.sync()
    print("iter", it, "routed", n, "of", r.positions, "hits", h.n)
    for kname, (ms, nl) in sorted(ctx.prof_report().items()):
        print("   %-22s %9.4f ms  x%d" % (kname, ms, nl))
    h.free(); r.free()
