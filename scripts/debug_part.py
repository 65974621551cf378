"""debug helper: route / lookup / collect once on a tiny input (run under compute-sanitizer)"""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from superplus_b200 import api, synth
inp = synth.make_config("tiny")
k, n_part = 25, int(sys.argv[1]) if len(sys.argv) > 1 else 2
ctx = api.Context(0)
rs, cs = ctx.upload(inp.reads[:6]), ctx.upload(inp.contigs)
r = ctx.route_plan(rs, k, n_part, 0, rs.tiles)
print("counts", r.counts, r.kmers)
n = r.kmers
keys = torch.empty(n, dtype=torch.int64, device="cuda"); ans = torch.empty(n, dtype=torch.int64, device="cuda")
r.keys(keys.data_ptr()); ctx.sync(); print("keys ok")
c = ctx.route_plan(cs, k, n_part, 0, cs.tiles)
recs = torch.empty(2 * c.kmers, dtype=torch.int64, device="cuda")
c.records(recs.data_ptr()); ctx.sync(); print("records ok")
t = ctx.table_create(c.kmers, k); t.insert_records(recs.data_ptr(), c.kmers); ctx.sync(); print("insert ok", t.stats())
t.lookup_keys(keys.data_ptr(), n, ans.data_ptr()); ctx.sync(); print("lookup ok")
h = r.collect(ans.data_ptr()); print("collect ok", h.n)
