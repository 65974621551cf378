import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from superplus_b200 import api, synth
from oracle import oracle as orc
O = orc.Oracle()
ctx = api.Context(0)
P = api.make_sw_params(); OP = orc.make_params()
for (ql, tl, n) in [(2000, 500, 4), (2000, 1000, 4), (4000, 1500, 4), (6000, 1800, 4), (10000, 1900, 4), (10000, 2000, 6), (10000, 2000, 2), (3000, 2000, 3)]:
    q, t = synth.make_sw_pairs(n, ql, tl, seed=46)
    want = [O.sw_align(OP, q[p], t[p], 1) for p in range(n)]
    for force in (None, "1"):
        if force: os.environ["GCG_SW_FORCE_GENERIC"] = "1"
        elif "GCG_SW_FORCE_GENERIC" in os.environ: del os.environ["GCG_SW_FORCE_GENERIC"]
        for rep in range(2):
            res, cigs = ctx.sw_batch(P, list(q), list(t), 1)
            bad = [p for p in range(n) if (int(res[p]["score"]), int(res[p]["bt_tidx"]), int(res[p]["bt_qidx"]), orc.cigar_str(cigs[p])) != (want[p]["score"], want[p]["bt_tidx"], want[p]["bt_qidx"], orc.cigar_str(want[p]["cigar"]))]
            print(ql, tl, n, "generic" if force else "packed", "rep", rep, "bad pairs", bad, [(int(res[p]["score"]), want[p]["score"]) for p in bad][:4], flush=True)
