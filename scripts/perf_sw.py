"""Quick SW timing on the GPU box (not the bench contract; see bench.py)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from superplus_b200 import api, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 1
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
qlen, tlen = int(os.environ.get("QLEN", 10_000)), int(os.environ.get("TLEN", 2_000))
t0 = time.time()
base_q, base_t = synth.make_sw_pairs(min(n, 256), qlen, tlen, seed=46)
reps = (n + len(base_q) - 1) // len(base_q)
q = np.tile(base_q, (reps, 1))[:n]; t = np.tile(base_t, (reps, 1))[:n]
print("gen %.1fs pairs=%d" % (time.time() - t0, n), flush=True)
ctx = api.Context(0)
b = ctx.swbatch_upload(q, t)
P = api.make_sw_params()
ctx.prof(True)
for it in range(iters):
    ctx.prof_reset()
    t0 = time.time()
    b.align(P, mode)
    dt = time.time() - t0
    rep = ctx.prof_report()
    cells = b.cells()
    print("iter %d wall %.4fs  %.1f GCUPS (wall)  paths %s" % (it, dt, cells / dt / 1e9, b.path_counts()))
    for kname, (ms, nl) in sorted(rep.items()):
        print("   %-22s %10.4f ms  x%d   %.1f GCUPS" % (kname, ms, nl, cells / ms / 1e6))
res, cigs = b.download()
print("scores", res["score"][:8], "n_cigar", res["n_cigar"][:8])
if len(sys.argv) > 4:
    os.environ["GCG_SW_FORCE_GENERIC"] = "1"
    m = min(n, 512)
    b2 = ctx.swbatch_upload(q[:m], t[:m])
    for it in range(2):
        ctx.prof_reset(); t0 = time.time(); b2.align(P, mode); dt = time.time() - t0
        print("generic: wall %.4fs %.1f GCUPS" % (dt, b2.cells() / dt / 1e9), ctx.prof_report())
    r2, c2 = b2.download()
    assert np.array_equal(r2["score"], res["score"][:m])
