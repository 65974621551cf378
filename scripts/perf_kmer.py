"""Quick k-mer path timing on the GPU box (not the bench contract; see bench.py)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from superplus_b200 import api, synth

name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
k = int(sys.argv[2]) if len(sys.argv) > 2 else 25
t0 = time.time()
inp = synth.make_config(name)
print("gen %.1fs reads=%d bases=%d contigs=%d" % (time.time() - t0, len(inp.reads), sum(len(r) for r in inp.reads), len(inp.contigs)), flush=True)
ctx = api.Context(0, host_threads=8)
contigs = inp.contigs
t0 = time.time(); cs = ctx.upload(contigs); t1 = time.time(); rs = ctx.upload(inp.reads); t2 = time.time()
print("upload contigs %.3fs reads %.3fs" % (t1 - t0, t2 - t1), flush=True)
ctx.prof(True)
for it in range(int(sys.argv[3]) if len(sys.argv) > 3 else 4):
    ctx.prof_reset()
    t0 = time.time()
    t = ctx.table_build(cs, k)
    n = ctx.search_device_compact(t, rs) if os.environ.get('GCG_SEARCH_FUSED', '1') != '0' else ctx.search_device(t, rs)
    st = t.stats()
    dt = time.time() - t0
    rep = ctx.prof_report()
    print("iter", it, "wall %.4fs" % dt, "hits", n, "stats", st)
    for kname, (ms, nl) in sorted(rep.items()):
        print("   %-18s %9.4f ms  x%d" % (kname, ms, nl))
    nk = rs.kmers(k)
    if "k45_fused" in rep:
        ms = rep["k45_fused"][0]
        print("   K4+K5 fused: %.1f G k-mers/s ; algorithmic %.1f GB/s" % (nk / ms / 1e6, (16.25 * nk + 16 * n) / ms / 1e6))
    if "k45_search" in rep:
        ms = rep["k45_search"][0] + rep.get("hits_emit", (0, 0))[0] + sum(rep.get(x, (0, 0))[0] for x in ("scan_reduce", "scan_blocksums", "scan_apply"))
        print("   K4+K5 two-pass (probe + scan + emit): %.1f G k-mers/s ; algorithmic %.1f GB/s" % (nk / ms / 1e6, (16.25 * nk + 16 * n) / ms / 1e6))
    if "k23_build" in rep:
        ms = rep["k23_build"][0]
        print("   build : %.2f G k-mers/s ; algorithmic %.1f GB/s" % (cs.kmers(k) / ms / 1e6, 32.25 * cs.kmers(k) / ms / 1e6))
    t.free()
t0 = time.time()
tb = ctx.table_build(cs, k)
hits = ctx.search_host(tb, inp.reads)
print("e2e host search (16-byte anchors) %.3fs hits %d" % (time.time() - t0, len(hits)))
for it in range(3):
    tb2 = ctx.table_build(cs, k)
    ctx.prof_reset()
    t0 = time.time()
    a, ro = ctx.search_host_compact(tb2, inp.reads)
    print("e2e host search (compact anchors) %.4fs hits %d" % (time.time() - t0, len(a)), {n_: (round(v[0], 3), v[1]) for n_, v in ctx.prof_report().items()})
    tb2.free()
words, woff, lens, keep = api.Context.pack_reads(inp.reads, pinned=True)
import ctypes as C
for it in range(3):
    tb2 = ctx.table_build(cs, k)
    ctx.prof_reset()
    ap, rp, na = C.c_void_p(), C.c_void_p(), C.c_int64()
    t0 = time.time()
    ctx._chk(ctx.L.gcg_search_compact_packed(ctx.h, tb2.h, words.ctypes.data, woff.ctypes.data, lens.ctypes.data, len(lens), k, C.byref(ap), C.byref(rp), C.byref(na)))
    dt = time.time() - t0
    ctx.L.gcg_free(ap); ctx.L.gcg_free(rp)
    print("e2e packed search %.4fs hits %d" % (dt, na.value), {n_: (round(v[0], 3), v[1]) for n_, v in ctx.prof_report().items()})
    tb2.free()
