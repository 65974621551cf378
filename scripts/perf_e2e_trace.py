"""the bench's end-to-end k-mer step (gcg_table_build + gcg_search_compact[_packed] + gcg_table_stats) call by call,
with the library's phase trace (GCG_TRACE=1) of the last iteration on stderr"""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("GCG_TRACE", "1")
import numpy as np
from superplus_b200 import api, synth
cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
nt = int(sys.argv[2]) if len(sys.argv) > 2 else min(16, os.cpu_count())
K = 25
inp = synth.make_config(cfg)
ctx = api.Context(0, host_threads=nt)
arrs = [np.ascontiguousarray(r) for r in inp.reads]
rptrs = (C.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])
rlens = np.array([len(a) for a in arrs], dtype=np.int32)
carrs = [np.ascontiguousarray(c) for c in inp.contigs]
cptrs = (C.c_void_p * len(carrs))(*[a.ctypes.data for a in carrs])
clens = np.array([len(a) for a in carrs], dtype=np.int32)
pk_words, pk_woff, pk_lens, pk_keep = api.Context.pack_reads(arrs, pinned=True)
n_kmers = sum(max(0, len(a) - K + 1) for a in arrs)
for mode in ("ascii", "packed"):
    for it in range(6):
        sys.stderr.write("---- %s iteration %d\n" % (mode, it)); sys.stderr.flush()
        t0 = time.perf_counter()
        h = C.c_void_p()
        ctx._chk(ctx.L.gcg_table_build(ctx.h, C.cast(cptrs, C.c_void_p), clens.ctypes.data, len(carrs), K, C.byref(h)))
        tab = api.KmerTable(ctx, h, K)
        t1 = time.perf_counter()
        hp, rp, nh = C.c_void_p(), C.c_void_p(), C.c_int64()
        if mode == "ascii":
            ctx._chk(ctx.L.gcg_search_compact(ctx.h, tab.h, C.cast(rptrs, C.c_void_p), rlens.ctypes.data, len(arrs), K, C.byref(hp), C.byref(rp), C.byref(nh)))
        else:
            ctx._chk(ctx.L.gcg_search_compact_packed(ctx.h, tab.h, pk_words.ctypes.data, pk_woff.ctypes.data, pk_lens.ctypes.data, len(arrs), K, C.byref(hp), C.byref(rp), C.byref(nh)))
        t2 = time.perf_counter()
        ctx.L.gcg_free(hp); ctx.L.gcg_free(rp)
        t3 = time.perf_counter()
        s4 = tab.stats()
        t4 = time.perf_counter()
        tab.free()
        t5 = time.perf_counter()
        print("%s it %d: build %.3f search %.3f free %.3f stats %.3f tfree %.3f total %.3f ms = %.1f G k-mers/s" %
              (mode, it, 1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2), 1e3 * (t4 - t3), 1e3 * (t5 - t4), 1e3 * (t5 - t0), n_kmers / (t5 - t0) / 1e9), flush=True)
