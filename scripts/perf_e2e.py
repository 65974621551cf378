"""e2e host-buffer search timing with the phase trace (GCG_TRACE=1) for several host thread counts"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from superplus_b200 import api, synth
inp = synth.make_config(sys.argv[1] if len(sys.argv) > 1 else "cfg2")
k = 25
for nt in [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "4,8,16").split(",")]:
    ctx = api.Context(0, host_threads=nt)
    cs = ctx.upload(inp.contigs)
    tb = ctx.table_build(cs, k)
    for it in range(3):
        t0 = time.perf_counter()
        hits = ctx.search_host(tb, inp.reads)
        print("threads %d iter %d: %.2f ms (incl. %d-anchor numpy copy)" % (nt, it, 1e3 * (time.perf_counter() - t0), len(hits)), flush=True)
    tb.free(); cs.free(); ctx.close()
